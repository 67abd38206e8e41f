#!/usr/bin/env python
"""Pipeline experiment: all HBM sweeps (K1, K3) of all scales and steps through ONE stream (strictly one bandwidth
kernel at a time, K1 of step i ahead of K3 of step i - SKEW), the morphology kernels (K2) on high-priority streams
in the background.  Compare with bench.py's scheme (one stream per scale, 4 steps in flight).
    python tools/bw_stream_probe.py [--skew 2] [--steps 400]"""
import argparse, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from golden_util import weights
from mcaq_yolo_b200 import modules as M, ops, constants as K
from mcaq_yolo_b200.fused import mapper_block, ScaleWorkspace

ap = argparse.ArgumentParser()
ap.add_argument("--skew", type=int, default=2)
ap.add_argument("--steps", type=int, default=400)
ap.add_argument("--k2-streams", type=int, default=2)
ap.add_argument("--batch", type=int, default=64)
ap.add_argument("--k1-parallel", action="store_true", help="K1 of the three scales on three streams instead of one")
a_ = ap.parse_args()
dev = torch.device("cuda")
B = a_.batch
W = weights()
an, mp, _ = M.build_fixture_modules(W, device=dev)
shapes = [(64, 80, 80), (128, 40, 40), (256, 20, 20)]
qs = [M.build_fixture_modules(W, device=dev)[2] for _ in shapes]
g = torch.Generator(device=dev); g.manual_seed(1234)
NS = 4
sets = []
for _ in range(NS):
    feats = []
    for C, H, Wd in shapes:
        coarse = torch.randn(B, C, H // 8 + 2, Wd // 8 + 2, device=dev, generator=g)
        up = torch.nn.functional.interpolate(coarse, size=(H, Wd), mode="bilinear", align_corners=False)
        feats.append((up * 1.6 + 0.1 * torch.randn(B, C, H, Wd, device=dev, generator=g) + 0.3).bfloat16().contiguous())
    sets.append(feats)
from mcaq_yolo_b200 import _lib
_lib.load().mcaq_morph_policy(0)
cm = K.pack_complexity_mlp(an.complexity_mlp)
mb = mapper_block(mp, 1.0)
sms = [K.pack_soft_mask(q.soft_mask) for q in qs]
slots = []
for j in range(NS):
    sl = {"ws": [ScaleWorkspace(C, dev) for C, _, _ in shapes],
          "s": [torch.empty((B, H, Wd), device=dev) for _, H, Wd in shapes],
          "a": [torch.empty((B, H, Wd), device=dev) for _, H, Wd in shapes], "r": [None] * 3, "y": [None] * 3}
    slots.append(sl)
bw = torch.cuda.Stream()
k2s = [torch.cuda.Stream(priority=-1) for _ in range(a_.k2_streams)]
k2side = [[torch.cuda.Stream(priority=-1) for _ in range(2)] for _ in range(a_.k2_streams)]

def phaseA(j):
    sl = slots[j]
    for i, x in enumerate(sets[j]):
        ops.reduce_planes_into(x, sl["s"][i], sl["a"][i], sl["ws"][i].keys)

def phaseK(j, q):
    sl = slots[j]
    cur = torch.cuda.current_stream()
    fork = torch.cuda.Event(); fork.record(cur)
    joins = []
    for i, (C, H, Wd) in enumerate(shapes):
        st = cur if i == 0 else k2side[q][i - 1]
        if i:
            st.wait_event(fork)
        with torch.cuda.stream(st):
            sl["r"][i] = ops.morph_fused(sl["s"][i], sl["a"][i], C, an.grid_size, cm, mb, sms[i], 1.0, False, sl["ws"][i].keys,
                                         mp.min_bits, mp.max_bits, 1e-3)
            if i:
                ev = torch.cuda.Event(); ev.record(st); joins.append(ev)
    for ev in joins:
        cur.wait_event(ev)

def phaseB(j):
    sl = slots[j]
    for i, x in enumerate(sets[j]):
        r = sl["r"][i]
        sl["y"][i] = ops.tile_quantize_ranges(x, r["bit_map"], r["packed"], None, None, r["mask"])

with torch.no_grad():
    for j in range(NS):                                   # eager warm-up (also builds the step table etc.)
        phaseA(j); phaseK(j, 0); phaseB(j)
    torch.cuda.synchronize()
    GA, GK, GB = [], [], []
    for j in range(NS):
        ga = torch.cuda.CUDAGraph()
        with torch.cuda.graph(ga, stream=bw):
            phaseA(j)
        gk = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gk, stream=k2s[j % a_.k2_streams]):
            phaseK(j, j % a_.k2_streams)
        gb = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gb, stream=bw):
            phaseB(j)
        GA.append(ga); GK.append(gk); GB.append(gb)
    torch.cuda.synchronize()
    evA = [torch.cuda.Event() for _ in range(NS)]
    evK = [torch.cuda.Event() for _ in range(NS)]
    evB = [torch.cuda.Event() for _ in range(NS)]

    def run(n):
        for i in range(n + a_.skew):
            if i < n:
                j = i % NS
                with torch.cuda.stream(bw):
                    GA[j].replay()
                    evA[j].record(bw)
                kq = k2s[j % a_.k2_streams]
                kq.wait_event(evA[j])
                with torch.cuda.stream(kq):
                    GK[j].replay()
                    evK[j].record(kq)
            ib = i - a_.skew
            if ib >= 0:
                jj = ib % NS
                bw.wait_event(evK[jj])
                with torch.cuda.stream(bw):
                    GB[jj].replay()
    run(40)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(bw)
    run(a_.steps)
    e1.record(bw)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a_.steps
    alg = 3 * 2 * sum(C * H * Wd for C, H, Wd in shapes) * B
    print("bw-stream pipeline: skew %d, %d K2 streams: %.4f ms/step, %.0f images/s, %.3f of 6541.8 GB/s" %
          (a_.skew, a_.k2_streams, ms, B / ms * 1e3, alg / ms / 1e6 / 6541.8))
