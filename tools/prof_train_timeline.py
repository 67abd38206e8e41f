#!/usr/bin/env python
"""Kernel timeline of ONE replay of the captured training step (one stream per scale): per-stream busy time, the
critical path's idle gaps and the longest kernels.  python tools/prof_train_timeline.py [one|scale]"""
import os, sys, collections
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from mcaq_yolo_b200 import modules as M
from mcaq_yolo_b200 import train_nets as TN
from golden_util import weights
mode = sys.argv[1] if len(sys.argv) > 1 else "scale"
W = weights()
B = 16
shapes = [(64, 80, 80), (128, 40, 40), (256, 20, 20)]
a, m, _ = M.build_fixture_modules(W, "cuda"); a.train(); m.train()
qs = [M.build_fixture_modules(W, "cuda")[2].train() for _ in shapes]
params = [p for mod in [a, m] + qs for p in mod.parameters()]
feats = [(torch.randn(B, C, H, Wd, device="cuda") * 2 + 0.3).bfloat16() for C, H, Wd in shapes]
teach = [torch.randn(B, C, H, Wd, device="cuda") for C, H, Wd in shapes]
gouts = [torch.randn(B, C, H, Wd, device="cuda").bfloat16() * 1e-3 for C, H, Wd in shapes]
SS = [torch.cuda.Stream() for _ in shapes]
def step():
    for p in params:
        p.grad = None
    main = torch.cuda.current_stream()
    if mode == "scale" and os.environ.get("NO_PREPARE") != "1":
        TN.prepare_step(a.complexity_mlp, m.mapping_network)
    parts, bits = [], []
    for si, (x0, t, go, q) in enumerate(zip(feats, teach, gouts, qs)):
        st = SS[si] if mode == "scale" else main
        if mode == "scale":
            st.wait_stream(main)
        with torch.cuda.stream(st):
            x = x0.detach().requires_grad_(True)
            q.kd_teacher = t
            r = M.mcaq_hook_forward(x, a, m, q, temperature=1.0, training=True)
            part = (r["features_q"] * go).sum().float() + r["kd_feature_loss"] / 3
        parts.append(part); bits.append(r["bit_map"])
    if mode == "scale":
        for st in SS:
            main.wait_stream(st)
    _, lbit, _ = TN.bit_map_losses(bits, 4.0)
    (torch.stack(parts).sum() + 0.1 * lbit).backward()
side = torch.cuda.Stream(); side.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(side):
    for _ in range(3):
        step()
torch.cuda.current_stream().wait_stream(side); torch.cuda.synchronize()
for p in params:
    p.grad = None
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    step()
for _ in range(3):
    g.replay()
torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    g.replay()
    torch.cuda.synchronize()
ev = [e for e in prof.events() if e.device_type.name == "CUDA" and e.time_range.end > e.time_range.start]
t0 = min(e.time_range.start for e in ev)
rows = sorted((e.time_range.start - t0, e.time_range.end - t0, getattr(e, "device_resource_id", -1), e.name[:60]) for e in ev)
print("kernels", len(rows), "span us", max(r[1] for r in rows))
busy = collections.defaultdict(float)
for s, e, st, n in rows:
    busy[st] += e - s
print("busy per stream", dict(busy))
# union of busy intervals = GPU busy; gaps = nothing running
cur_e, idle = 0.0, 0.0
for s, e, st, n in rows:
    if s > cur_e:
        idle += s - cur_e
    cur_e = max(cur_e, e)
print("idle (no kernel running) us", round(idle, 1))
for s, e, st, n in rows:
    print(f"{s:8.1f} {e - s:7.1f}  s{st}  {n}")
