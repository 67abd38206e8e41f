#!/usr/bin/env python
"""Which parts of the training hook survive CUDA-graph capture?  Captures each stage on its own and
reports the first failure (debug aid for tools/train_bench.py's captured training step)."""
import os
import sys
import traceback

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
from golden_util import weights  # noqa: E402
from mcaq_yolo_b200 import modules as M  # noqa: E402
from mcaq_yolo_b200 import ops  # noqa: E402

dev = "cuda"
W = weights()
a, m, q = M.build_fixture_modules(W, device=dev)
a.train(); m.train(); q.train()
x = (torch.randn(4, 64, 80, 80, device=dev) * 2 + 0.3)
t = torch.randn_like(x)


def stage(name, fn, warm=2):
    try:
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            for _ in range(warm):
                fn()
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            out = fn()
        g.replay()
        torch.cuda.synchronize()
        print(f"[ok]   {name}")
        return out
    except Exception as e:      # noqa: BLE001
        print(f"[FAIL] {name}: {type(e).__name__}: {str(e)[:300]}")
        traceback.print_exc(limit=6)
        torch.cuda.synchronize()
        return None


stage("K1 reduce_planes", lambda: ops.reduce_planes(x))
stage("phi", lambda: a.compute_phi_tiles(x))
cm = stage("analyzer train forward", lambda: a(x))
if cm is None:
    cm = a(x)
bits = stage("mapper train forward", lambda: m(cm.detach(), 1.0, return_continuous=True))
if bits is None:
    bits = m(cm.detach(), 1.0, return_continuous=True)
stage("update_running_stats", lambda: q.update_running_stats(x))
stage("soft mask (torch path)", lambda: q.soft_mask(bits.detach().requires_grad_(True), x))
stage("quantizer train forward", lambda: q(x, bits.detach(), training=True))


def fwd_bwd():
    for p in list(a.parameters()) + list(m.parameters()) + list(q.parameters()):
        p.grad = None
    xx = x.detach().requires_grad_(True)
    q.kd_teacher = t
    r = M.mcaq_hook_forward(xx, a, m, q, temperature=1.0, training=True)
    loss = (r["features_q"] * 1e-3).sum() + r["kd_feature_loss"] + 0.1 * (r["bit_map"].mean() - 4.0) ** 2
    loss.backward()
    return loss


stage("hook forward + backward", fwd_bwd, warm=3)
