#!/usr/bin/env python
"""Driver for ncu captures of the fused K2 kernel on one shape: python tools/prof_k2.py C H [B]"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from mcaq_yolo_b200 import ops, _lib, constants as K, modules as M
_lib.load().mcaq_debug_cluster_split(int(os.environ.get('MCAQ_K2_SPLIT', '0')))
from golden_util import weights
C, H = int(sys.argv[1]), int(sys.argv[2])
B = int(sys.argv[3]) if len(sys.argv) > 3 else 64
a, m, q = M.build_fixture_modules(weights(), "cuda")
cm, mp, sm = K.pack_complexity_mlp(a.complexity_mlp), K.pack_mapping_steps(m.mapping_network, 1.0, m.min_bits, m.max_bits), K.pack_soft_mask(q.soft_mask)
coarse = torch.randn(B, C, H // 8 + 2, H // 8 + 2, device="cuda")
x = (torch.nn.functional.interpolate(coarse, size=(H, H), mode="bilinear") * 1.6 + 0.1 * torch.randn(B, C, H, H, device="cuda")).to(torch.bfloat16)
s, ab, k = ops.reduce_planes(x)
for _ in range(4):
    r = ops.morph_fused(s, ab, C, 8, cm, mp, sm, 1.0)
torch.cuda.synchronize()
print("ok", r["bit_map"].float().mean().item())
