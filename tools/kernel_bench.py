#!/usr/bin/env python
"""Per-kernel timing in isolation (CUDA events, rotating buffers larger than L2) for the
bandwidth kernels.  Usage: python tools/kernel_bench.py [--dtype bf16|f32] [--batch 64] [--iters 30]"""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mcaq_yolo_b200 import ops  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--dtype", default="bf16")
ap.add_argument("--batch", type=int, default=64)
ap.add_argument("--iters", type=int, default=30)
ap.add_argument("--shapes", default="64x80x80,128x40x40,256x20x20")
ap.add_argument("--only", default="")
a = ap.parse_args()
dt = torch.bfloat16 if a.dtype == "bf16" else torch.float32
es = 2 if a.dtype == "bf16" else 4
dev = "cuda"
B = a.batch


def timeit(fn, nbuf, iters):
    """Median / min time per launch from CUDA-graph replays of `nbuf` launches over rotating
    buffers (a replay keeps the GPU queue full, so host launch latency is not measured)."""
    for i in range(nbuf):
        fn(i)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for i in range(nbuf):
            fn(i)
    g.replay()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        g.replay()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3 / nbuf)
    ts.sort()
    return ts[len(ts) // 2], ts[0]


for shp in a.shapes.split(","):
    C, H, W = [int(v) for v in shp.split("x")]
    nbytes = B * C * H * W * es
    nbuf = max(2, int(400e6 // nbytes) + 1)
    xs = [(torch.randn(B, C, H, W, device=dev) * 2 + 0.3).to(dt) for _ in range(nbuf)]
    ys = [torch.empty_like(x) for x in xs]
    tile = ops.tile_size(H, 8)
    Ht, Wt = H // tile, W // tile
    bm = torch.randint(2, 9, (B, Ht, Wt), device=dev).float()
    m = torch.rand(B, H, W, device=dev) * 0.2 + 0.8
    s, ab, keys = ops.reduce_planes(xs[0])
    qt = ops.build_qtable(ops.ranges_decode(keys))
    if a.only in ("", "k1"):
        med, mn = timeit(lambda i: ops.reduce_planes(xs[i]), nbuf, a.iters)
        print(f"K1 reduce_planes  {shp:12s} {a.dtype} B={B}: median {med:7.1f} us (min {mn:7.1f})  "
              f"{nbytes / med / 1e3:7.0f} GB/s   [{nbuf} rotating bufs]")
        med, mn = timeit(lambda i: ops.reduce_planes(xs[i], want_ranges=False), nbuf, a.iters)
        print(f"K1 (no ranges)    {shp:12s} {a.dtype} B={B}: median {med:7.1f} us (min {mn:7.1f})  "
              f"{nbytes / med / 1e3:7.0f} GB/s")
    if a.only in ("", "k3"):
        med, mn = timeit(lambda i: ops.tile_quantize(xs[i], bm, qt, m, out=ys[i]), nbuf, a.iters)
        print(f"K3 tile_quantize  {shp:12s} {a.dtype} B={B}: median {med:7.1f} us (min {mn:7.1f})  "
              f"{2 * nbytes / med / 1e3:7.0f} GB/s")
        med, mn = timeit(lambda i: ops.tile_quantize(xs[i], bm, qt, None, out=ys[i]), nbuf, a.iters)
        print(f"K3 (no mask)      {shp:12s} {a.dtype} B={B}: median {med:7.1f} us (min {mn:7.1f})  "
              f"{2 * nbytes / med / 1e3:7.0f} GB/s")
        pk = ops.ranges_decode(keys)
        for tag, flag in (("K3 ranges (LDG) ", False), ("K3 ranges (TMA) ", True)):
            ops.K3_TMA = flag
            try:
                med, mn = timeit(lambda i: ops.tile_quantize_ranges(xs[i], bm, pk, None, None, m, out=ys[i]), nbuf, a.iters)
                print(f"{tag}  {shp:12s} {a.dtype} B={B}: median {med:7.1f} us (min {mn:7.1f})  "
                      f"{2 * nbytes / med / 1e3:7.0f} GB/s")
            except RuntimeError as e:
                print(f"{tag}  {shp:12s}: {e}")
        ops.K3_TMA = False
        med, mn = timeit(lambda i: ys[i].copy_(xs[i]), nbuf, a.iters)
        print(f"torch copy_       {shp:12s} {a.dtype} B={B}: median {med:7.1f} us (min {mn:7.1f})  "
              f"{2 * nbytes / med / 1e3:7.0f} GB/s")
