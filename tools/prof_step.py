#!/usr/bin/env python
"""Driver for ncu captures: a few eager steps of the fused hot path on the bench workload
(BASELINE configs[1]: v8n@640, batch 64, bf16; or `fp32`), one stream, 9 kernels per step.
  ncu --set full -k regex:'reduce_planes|morph_fused|tile_quantize' --launch-skip 18 --launch-count 9 \\
      python tools/prof_step.py [bf16|f32] [batch]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from golden_util import weights  # noqa: E402
from mcaq_yolo_b200 import modules as M  # noqa: E402
from mcaq_yolo_b200.fused import FusedHotPath  # noqa: E402

dt = torch.float32 if len(sys.argv) > 1 and sys.argv[1] == "f32" else torch.bfloat16
B = int(sys.argv[2]) if len(sys.argv) > 2 else 64
dev = torch.device("cuda")
W = weights()
a, m, _ = M.build_fixture_modules(W, device=dev)
shapes = [(64, 80, 80), (128, 40, 40), (256, 20, 20)]
qs = [M.build_fixture_modules(W, device=dev)[2] for _ in shapes]
g = torch.Generator(device=dev)
g.manual_seed(1234)
feats = []
for C, H, Wd in shapes:
    coarse = torch.randn(B, C, H // 8 + 2, Wd // 8 + 2, device=dev, generator=g)
    up = torch.nn.functional.interpolate(coarse, size=(H, Wd), mode="bilinear", align_corners=False)
    feats.append((up * 1.6 + 0.1 * torch.randn(B, C, H, Wd, device=dev, generator=g) + 0.3).to(dt).contiguous())
hot = FusedHotPath(a, m, qs, streams=False)
with torch.no_grad():
    for _ in range(4):
        out = hot.run(feats)
torch.cuda.synchronize()
print("ok", float(out[0]["bit_map"].mean()))
