#!/usr/bin/env python
"""Where the small torch kernels of one training step come from: torch.profiler with Python stacks, grouped by
(op, innermost repo frame).  python tools/prof_train_ops.py"""
import os, sys, collections
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from mcaq_yolo_b200 import modules as M
from mcaq_yolo_b200 import train_nets as TN
from golden_util import weights
W = weights()
B = 16
shapes = [(64, 80, 80), (128, 40, 40), (256, 20, 20)]
a, m, _ = M.build_fixture_modules(W, "cuda"); a.train(); m.train()
qs = [M.build_fixture_modules(W, "cuda")[2].train() for _ in shapes]
params = [p for mod in [a, m] + qs for p in mod.parameters()]
feats = [(torch.randn(B, C, H, Wd, device="cuda") * 2 + 0.3).bfloat16() for C, H, Wd in shapes]
teach = [torch.randn(B, C, H, Wd, device="cuda") for C, H, Wd in shapes]
def step():
    for p in params:
        p.grad = None
    parts, bits = [], []
    for x0, t, q in zip(feats, teach, qs):
        x = x0.detach().requires_grad_(True)
        q.kd_teacher = t
        r = M.mcaq_hook_forward(x, a, m, q, temperature=1.0, training=True)
        parts.append((r["features_q"].float() * 1e-3).sum() + r["kd_feature_loss"] / 3)
        bits.append(r["bit_map"])
    _, lbit, _ = TN.bit_map_losses(bits, 4.0)
    (torch.stack(parts).sum() + 0.1 * lbit).backward()
for _ in range(3):
    step()
torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], with_stack=True) as prof:
    step()
    torch.cuda.synchronize()
cnt = collections.Counter()
for e in prof.events():
    if e.device_type.name != "CPU" or not e.name.startswith("aten::"):
        continue
    if not getattr(e, "kernels", None):                    # only ops that launched a kernel themselves
        continue
    frame = next((s for s in (e.stack or []) if "/repo/" in s and "prof_train_ops" not in s), None) or \
            next((s for s in (e.stack or []) if "prof_train_ops" in s), "autograd engine / no python frame")
    cnt[(e.name, frame.split("/repo/")[-1][:110])] += 1
for (name, frame), n in cnt.most_common(60):
    print(f"{n:4d}  {name:28s} {frame}")
