#!/usr/bin/env python
"""Per-stage clock64() breakdown of the fused morphology kernel (debug aid)."""
import ctypes, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from mcaq_yolo_b200 import ops, _lib, constants as K, modules as M
from golden_util import weights
lib = _lib.load()
names = ["S0-1 load/norm", "S3 adaptive", "S4 lbp+sobel", "S5 phi2/3", "S6 blur", "S7|S8 otsu|mag", "S9 nms",
         "S10 hyst", "S11 counts", "S11b boxes", "S12 phi", "N1 complexity", "N2 mapper", "N3 softmask"]
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
a, m, q = M.build_fixture_modules(weights(), "cuda")
cm, mp, sm = K.pack_complexity_mlp(a.complexity_mlp), K.pack_mapping_network(m.mapping_network), K.pack_soft_mask(q.soft_mask)
for (C, H) in ((64, 80), (128, 40), (256, 20), (128, 160)):
    x = (torch.randn(B, C, H, H, device="cuda") * 2).to(torch.bfloat16)
    s, ab, k = ops.reduce_planes(x)
    clk = torch.zeros(B, 16, dtype=torch.int64, device="cuda")
    lib.mcaq_debug_stage_clocks(clk.data_ptr())
    for _ in range(3):
        ops.morph_fused(s, ab, C, 8, cm, mp, sm, 1.0)
    torch.cuda.synchronize()
    lib.mcaq_debug_stage_clocks(None)
    c = clk.cpu().numpy()
    d = (c[:, 1:15] - c[:, 0:14]).mean(0)
    tot = (c[:, 14] - c[:, 0]).mean()
    print(f"C={C} H={H}: total {tot:.0f} cycles = {tot / 1.965e3:.1f} us @1.965GHz")
    for i in range(14):
        print(f"   {names[i]:16s} {d[i]:9.0f} cyc  {100 * d[i] / tot:5.1f}%")
