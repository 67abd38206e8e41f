#!/usr/bin/env python
"""Batched curriculum scoring: analyzer.score_image on raw images -- the reference scores its dataset one
image at a time through a Python loop (utils/dataset.py:276-401 -> core/morphology.py:923-937) -- here K1 + the plane
pipeline (csrc/morph_planes.cu) on a whole batch.  Prints one JSON line per (size, batch).
  python tools/score_bench.py [--ref]     (--ref: also time the reference's own analyzer on the same GPU / on CPU)"""
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
from mcaq_yolo_b200 import modules as M  # noqa: E402
from golden_util import weights  # noqa: E402

W = weights()
a, _, _ = M.build_fixture_modules(W, "cuda")
ref_a = None
if "--ref" in sys.argv:
    from harness import ref_model
    if ref_model.load(with_model=False) is not None:
        from mcaq_yolo.core import morphology
        ref_a = morphology.MorphologicalComplexityAnalyzer(grid_size=8, device="cuda")
        ref_a.load_state_dict({k: torch.as_tensor(v) for k, v in W["analyzer"].items()})
        ref_a = ref_a.to("cuda").eval()
        ref_c = morphology.MorphologicalComplexityAnalyzer(grid_size=8, device="cpu")
        ref_c.load_state_dict({k: torch.as_tensor(v) for k, v in W["analyzer"].items()})
        ref_c.eval()


def timed(fn, n):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


for size, B in ((640, 1), (640, 16), (640, 64), (1280, 1), (1280, 16)):
    x = torch.rand(B, 3, size, size, device="cuda")
    with torch.no_grad():
        ms = timed(lambda: a.score_image(x), 10)
        row = {"size": size, "batch": B, "ms": ms, "images_per_s": B / ms * 1e3}
        if ref_a is not None:
            n = max(1, min(B, 4))
            ms_r = timed(lambda: [ref_a.score_image(x[i:i + 1]) for i in range(n)], 2) / n      # its loop is per image
            row.update(reference_cuda_ms_per_image=ms_r, speedup_vs_reference_cuda=ms_r / (ms / B))
            s_nat, s_ref = a.score_image(x[:n]), torch.cat([ref_a.score_image(x[i:i + 1]) for i in range(n)])
            row["max_abs_score_diff_vs_reference_cuda"] = float((s_nat - s_ref).abs().max())
            if size == 640:       # the parity bar is the reference's CPU path (its CUDA path flips Otsu / threshold decisions on
                                  # some noise images: tools/score_check.py)
                s_cpu = torch.cat([ref_c.score_image(x[i:i + 1].cpu()) for i in range(n)])
                row["max_abs_score_diff_vs_reference_cpu"] = float((s_nat.cpu() - s_cpu).abs().max())
    print(json.dumps(row), flush=True)
