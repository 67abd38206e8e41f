#!/usr/bin/env python
"""Per-stage clock64() breakdown of the morphology kernel (debug aid)."""
import ctypes, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mcaq_yolo_b200 import ops, _lib, constants as K
lib = _lib.load()
lib.mcaq_debug_stage_clocks.argtypes = [ctypes.c_void_p]
lib.mcaq_debug_stage_clocks.restype = None
names = ["S0-2 load/norm", "S3 adaptive", "zero hist", "S4 lbp+sobel", "S5 phi2/3", "S6 blur", "S7 otsu", "S8 mag",
         "S9 nms", "S10 hyst", "S11 counts", "S11b boxes", "S12 phi"]
B = 64
for (C, H) in ((64, 80), (128, 40), (256, 20), (128, 160)):
    x = (torch.randn(B, C, H, H, device="cuda") * 2).to(torch.bfloat16)
    s, a, k = ops.reduce_planes(x)
    clk = torch.zeros(B, 16, dtype=torch.int64, device="cuda")
    lib.mcaq_debug_stage_clocks(clk.data_ptr())
    for _ in range(3):
        ops.morph_phi(s, C, 8, K.device_constants("cuda"))
    torch.cuda.synchronize()
    c = clk.cpu().numpy()
    d = (c[:, 1:13] - c[:, 0:12]).mean(0)
    tot = (c[:, 12] - c[:, 0]).mean()
    print(f"C={C} H={H}: total {tot:.0f} cycles")
    for n, v in zip(names[1:], d[1:] if False else d[0:]):
        pass
    for i in range(12):
        print(f"   {names[i]:16s} {d[i]:9.0f} cyc  {100 * d[i] / tot:5.1f}%")
    lib.mcaq_debug_stage_clocks(None)
