#!/usr/bin/env python
"""Per-stage clock64() breakdown of the fused morphology kernel (debug aid)."""
import ctypes, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from mcaq_yolo_b200 import ops, _lib, constants as K, modules as M
from golden_util import weights
lib = _lib.load()
names = ["L load+minmax", "N normalise", "T1 blur|adapt|lbp|act", "T2 mag+sync", "T3 otsu+nms+sync", "T4 hysteresis",
         "T5 counts+boxes", "phi + weights", "N1 cmlp+gather", "N1 bilateral", "N2 mapper", "N3 softmask"]
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
lib.mcaq_debug_cluster_split(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
a, m, q = M.build_fixture_modules(weights(), "cuda")
cm, mp, sm = K.pack_complexity_mlp(a.complexity_mlp), K.pack_mapping_steps(m.mapping_network, 1.0, m.min_bits, m.max_bits), K.pack_soft_mask(q.soft_mask)
for (C, H) in ((64, 80), (128, 40), (256, 20), (128, 160)):
    x = torch.nn.functional.interpolate(torch.randn(B, C, H // 8, H // 8, device="cuda"), size=(H, H), mode="bicubic")
    x = (x + 0.1 * torch.randn_like(x)).to(torch.bfloat16)
    s, ab, k = ops.reduce_planes(x)
    clk = torch.zeros(B, 16, dtype=torch.int64, device="cuda")
    lib.mcaq_debug_stage_clocks(clk.data_ptr())
    for _ in range(3):
        ops.morph_fused(s, ab, C, 8, cm, mp, sm, 1.0)
    torch.cuda.synchronize()
    lib.mcaq_debug_stage_clocks(None)
    c = clk.cpu().numpy()
    d = (c[:, 1:13] - c[:, 0:12]).mean(0)
    tot = (c[:, 12] - c[:, 0]).mean()
    print(f"C={C} H={H}: total {tot:.0f} cycles = {tot / 1.965e3:.1f} us @1.965GHz")
    for i in range(12):
        print(f"   {names[i]:24s} {d[i]:9.0f} cyc  {100 * d[i] / tot:5.1f}%")
    if (c[:, 13:16] > 0).all():       # inside N1 (first tile group of warp 0): layer 1 | LN64 + park | layer 2 | rest
        n1 = [(c[:, 13] - c[:, 8]).mean(), (c[:, 14] - c[:, 13]).mean(), (c[:, 15] - c[:, 14]).mean(), (c[:, 9] - c[:, 15]).mean()]
        print("   N1 detail: layer1 %.0f | LN64+park %.0f | layer2 %.0f | LN32+layer3+later rounds+sync %.0f" % tuple(n1))
