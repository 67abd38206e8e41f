#!/usr/bin/env python
"""Driver for ncu captures: a few eager steps of the fused hot path on the bench workload
(BASELINE configs[1]: v8n@640, batch 64, bf16; or `fp32`), one stream, 9 kernels per step.
  ncu --set full -k regex:'reduce_planes|morph_fused|tile_quantize' --launch-skip 36 --launch-count 36 \\
      python tools/prof_step.py [bf16|f32] [batch]
Eight steps over four rotating input sets; the last four are the steady state (cold L2 at every step)."""
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
NSETS = 4          # rotating input sets (4 x 92 MB > 126 MB L2): every step starts L2-cold, as in bench.py
sets = []
for _ in range(NSETS):
    feats = []
    for C, H, Wd in shapes:
        coarse = torch.randn(B, C, H // 8 + 2, Wd // 8 + 2, device=dev, generator=g)
        up = torch.nn.functional.interpolate(coarse, size=(H, Wd), mode="bilinear", align_corners=False)
        feats.append((up * 1.6 + 0.1 * torch.randn(B, C, H, Wd, device=dev, generator=g) + 0.3).to(dt).contiguous())
    sets.append(feats)
hot = FusedHotPath(a, m, qs, streams=False)
keep = []
with torch.no_grad():
    for i in range(8):                 # steps 0..7; capture steps 4..7 with --launch-skip 36 --launch-count 36
        out = hot.run(sets[i % NSETS])
        keep.append(out)               # outputs stay allocated: y of a step is not recycled (and re-dirtied) by the next
torch.cuda.synchronize()
print("ok", float(out[0]["bit_map"].mean()))
