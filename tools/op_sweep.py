#!/usr/bin/env python
"""Standalone complexity+quantize op sweep (BASELINE configs[4]): C3/C4/C5 shapes of
YOLOv8n@640 and YOLOv8s@1280, grid 4/8/16, fp32 / bf16, NCHW / channels_last.  Per kernel: time
per launch (CUDA events around a graph replay over rotating buffers > L2) and algorithmic HBM
bandwidth vs the measured peak; per hook: K1 + K2 + K3 serial.  Writes one JSON line per cell."""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from golden_util import weights  # noqa: E402
from mcaq_yolo_b200 import _lib, constants as K, fused, modules as M, ops  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--out", default="")
ap.add_argument("--iters", type=int, default=10)
a_ = ap.parse_args()
dev = torch.device("cuda")
peak = 6541.8
pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
if os.path.exists(pk):
    peak = float(json.load(open(pk))["hbm_gbs"])
W = weights()
SETS = {"v8n@640": (64, [(64, 80), (128, 40), (256, 20)]), "v8s@1280": (32, [(128, 160), (256, 80), (512, 40)])}
_lib.load().mcaq_morph_policy(1)          # serial hooks: latency policy


def timeit(fn, nbuf, iters):
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
    return ts[len(ts) // 2]


rows = []
for model, (B, shapes) in SETS.items():
    for grid in (4, 8, 16):
        an, mp_, q = M.build_fixture_modules(W, device=dev, grid_size=grid)
        cm, mpk, smk = K.pack_complexity_mlp(an.complexity_mlp), K.pack_mapping_steps(mp_.mapping_network, 1.0, mp_.min_bits, mp_.max_bits), K.pack_soft_mask(q.soft_mask)
        for (C, H) in shapes:
            for dt in (torch.bfloat16, torch.float32):
                for layout in ("nchw", "nhwc"):
                    es = 2 if dt == torch.bfloat16 else 4
                    nbytes = B * C * H * H * es
                    nbuf = max(2, int(300e6 // nbytes) + 1)
                    xs = []
                    for _ in range(nbuf):
                        coarse = torch.randn(B, C, H // 8 + 2, H // 8 + 2, device=dev)
                        x = (torch.nn.functional.interpolate(coarse, size=(H, H), mode="bilinear") * 1.6
                             + 0.1 * torch.randn(B, C, H, H, device=dev) + 0.3).to(dt)
                        xs.append(x.contiguous(memory_format=torch.channels_last) if layout == "nhwc" else x.contiguous())
                    ys = [torch.empty_like(x) for x in xs]
                    ws = fused.ScaleWorkspace(C, dev)
                    sp = torch.empty((B, H, H), device=dev)
                    apn = torch.empty((B, H, H), device=dev)
                    ops.reduce_planes_into(xs[0], sp, apn, ws.keys)
                    try:
                        r = ops.morph_fused(sp, apn, C, grid, cm, mpk, smk, 1.0, keys=ws.keys)
                    except RuntimeError as e:
                        rows.append({"model": model, "grid": grid, "C": C, "H": H, "dtype": str(dt)[6:], "layout": layout,
                                     "error": str(e)})
                        continue
                    t1 = timeit(lambda i: ops.reduce_planes_into(xs[i], sp, apn, ws.keys), nbuf, a_.iters)
                    t2 = timeit(lambda i: ops.morph_fused(sp, apn, C, grid, cm, mpk, smk, 1.0, keys=ws.keys), 4, a_.iters)
                    t3 = timeit(lambda i: ops.tile_quantize_ranges(xs[i], r["bit_map"], r["packed"], None, None, r["mask"],
                                                                   out=ys[i]), nbuf, a_.iters)
                    row = {"model": model, "batch": B, "grid": grid, "C": C, "H": H, "dtype": str(dt)[6:], "layout": layout,
                           "K1_us": round(t1, 1), "K2_us": round(t2, 1), "K3_us": round(t3, 1),
                           "K1_GBs": round(nbytes / t1 / 1e3), "K3_GBs": round(2 * nbytes / t3 / 1e3),
                           "K1_frac": round(nbytes / t1 / 1e3 / peak, 3), "K3_frac": round(2 * nbytes / t3 / 1e3 / peak, 3),
                           "hook_us": round(t1 + t2 + t3, 1),
                           "hook_frac": round(3 * nbytes / (t1 + t2 + t3) / 1e3 / peak, 3)}
                    rows.append(row)
                    print(json.dumps(row), flush=True)
                    del xs, ys
                    torch.cuda.empty_cache()
if a_.out:
    with open(a_.out, "w") as f:
        for r in rows:
            f.write(json.dumps(r) + "\n")
