#!/usr/bin/env python
"""Cycles warp 0 spends per T1 task type (adaptive / blur / LBP / activity) in the morphology kernel.
Needs a profiling build: MCAQ_NVCC_EXTRA="-DMCAQ_T1_PROF" python mcaq_yolo_b200/build.py --force"""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from mcaq_yolo_b200 import ops, _lib, constants as K, modules as M
from golden_util import weights
lib = _lib.load()
a, m, q = M.build_fixture_modules(weights(), "cuda")
cm, mp, sm = K.pack_complexity_mlp(a.complexity_mlp), K.pack_mapping_steps(m.mapping_network, 1.0, m.min_bits, m.max_bits), K.pack_soft_mask(q.soft_mask)
B = 64
for (C, H) in ((64, 80), (128, 40)):
    x = torch.nn.functional.interpolate(torch.randn(B, C, H // 8, H // 8, device="cuda"), size=(H, H), mode="bicubic")
    x = (x + 0.1 * torch.randn_like(x)).to(torch.bfloat16)
    s, ab, k = ops.reduce_planes(x)
    clk = torch.zeros(B, 16, dtype=torch.int64, device="cuda")
    lib.mcaq_debug_stage_clocks(clk.data_ptr())
    for _ in range(3):
        ops.morph_fused(s, ab, C, 8, cm, mp, sm, 1.0)
    torch.cuda.synchronize()
    lib.mcaq_debug_stage_clocks(None)
    c = clk.cpu().numpy().astype(float)
    print(f"C={C} H={H}: T1 stage {(c[:,3]-c[:,2]).mean():.0f} cyc; warp0: adaptive {c[:,13].mean():.0f} blur {c[:,14].mean():.0f} lbp {(c[:,15]//1000000).mean():.0f} act {(c[:,15]%1000000).mean():.0f}")
