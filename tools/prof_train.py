#!/usr/bin/env python
"""One eager training step of the three hooks (16 images, bf16) for `ncu --metrics gpu__time_duration.sum`:
python tools/prof_train.py"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from mcaq_yolo_b200 import modules as M
from golden_util import weights
W = weights()
B = 16
shapes = [(64, 80, 80), (128, 40, 40), (256, 20, 20)]
a, m, _ = M.build_fixture_modules(W, "cuda"); a.train(); m.train()
qs = [M.build_fixture_modules(W, "cuda")[2].train() for _ in shapes]
feats = [(torch.randn(B, C, H, Wd, device="cuda") * 2 + 0.3).bfloat16() for C, H, Wd in shapes]
teach = [torch.randn(B, C, H, Wd, device="cuda") for C, H, Wd in shapes]
params = [p for mod in [a, m] + qs for p in mod.parameters()]
def step():
    for p in params:                  # like optimizer.zero_grad(set_to_none=True): no accumulate-adds into stale grads
        p.grad = None
    loss = 0.0
    for x0, t, q in zip(feats, teach, qs):
        x = x0.detach().requires_grad_(True)
        q.kd_teacher = t
        r = M.mcaq_hook_forward(x, a, m, q, temperature=1.0, training=True)
        loss = loss + (r["features_q"].float() * 1e-3).sum() + r["kd_feature_loss"] / 3
    loss.backward()
for _ in range(3):
    step()
torch.cuda.synchronize()
print("MARK")
step()
torch.cuda.synchronize()
