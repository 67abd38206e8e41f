#!/usr/bin/env python
"""The reference's own CUDA kernel (oracle/_ref/libmcaq_ref_kernel.so, built by oracle/build_ref.py from
/root/reference/mcaq_yolo/ops/src/mcaq_kernel.cu) beside this library's Level-0 entry point
`launch_spatial_quantization`, on the C3 / C4 / C5 fp32 shapes of YOLOv8n@640 batch 64 and C3 of v8s@1280.
CUDA events around CUDA-graph replays over rotating buffers larger than L2.
  python tools/ref_kernel_bench.py [--out file.json]"""
import argparse
import ctypes
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
from mcaq_yolo_b200 import _lib, ops  # noqa: E402
import build_ref  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--out", default=None)
args = ap.parse_args()

ref_path = build_ref.build()
ref = None
if ref_path:
    dll = ctypes.CDLL(ref_path)
    ref = getattr(dll, build_ref.SYMBOL)
    ref.restype = None
    ref.argtypes = [ctypes.c_void_p] * 6 + [ctypes.c_int] * 8 + [ctypes.c_void_p]
lib = _lib.load()
peak = 6541.8
pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
if os.path.exists(pk):
    peak = float(json.load(open(pk))["hbm_gbs"])


def timed(fn, nbuf, reps=10, rounds=4):
    for i in range(nbuf):
        fn(i)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(rounds):
            for i in range(nbuf):
                fn(i)
    g.replay()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); g.replay(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) / (rounds * nbuf))
    ts.sort()
    return ts[len(ts) // 2]


rows = []
for name, B, C, H in (("C3 v8n@640", 64, 64, 80), ("C4 v8n@640", 64, 128, 40), ("C5 v8n@640", 64, 256, 20),
                      ("C3 v8s@1280", 32, 128, 160)):
    tile = ops.tile_size(H, 8)
    ht = H // tile
    nbytes = B * C * H * H * 4
    nbuf = max(2, int(160e6 // nbytes) + 1)
    xs = [torch.randn(B, C, H, H, device="cuda") * 2 + 0.3 for _ in range(nbuf)]
    ys = [torch.empty_like(x) for x in xs]
    bm = torch.randint(2, 9, (B, ht, ht), device="cuda").float()
    mask = torch.rand(B, 1, H, H, device="cuda") * 0.2 + 0.8
    mn = xs[0].amin(dim=(0, 2, 3)).contiguous()
    mx = xs[0].amax(dim=(0, 2, 3)).contiguous()
    st = lambda: torch.cuda.current_stream().cuda_stream     # noqa: E731

    def ours(i):
        lib.launch_spatial_quantization(xs[i].data_ptr(), bm.data_ptr(), mn.data_ptr(), mx.data_ptr(), mask.data_ptr(),
                                        ys[i].data_ptr(), B, C, H, H, tile, tile, ht, ht, st())

    def theirs(i):
        ref(xs[i].data_ptr(), bm.data_ptr(), mn.data_ptr(), mx.data_ptr(), mask.data_ptr(), ys[i].data_ptr(),
            B, C, H, H, tile, tile, ht, ht, st())

    t_ours = timed(ours, nbuf)
    assert lib.mcaq_level0_status() == 0
    y_ours = ys[0].clone()
    row = {"shape": name, "B": B, "C": C, "H": H, "bytes": 2 * nbytes, "ours_us": 1e3 * t_ours,
           "ours_GBs": 2 * nbytes / t_ours / 1e6, "ours_frac_of_peak": 2 * nbytes / t_ours / 1e6 / peak}
    if ref is not None:
        t_ref = timed(theirs, nbuf)
        y_ref = ys[0].clone()
        # the two differ only where the reference kernel's roundf (half away from zero) meets an exact .5
        diff = (y_ours != y_ref).float().mean().item()
        row.update(ref_us=1e3 * t_ref, ref_GBs=2 * nbytes / t_ref / 1e6, ref_frac_of_peak=2 * nbytes / t_ref / 1e6 / peak,
                   speedup=t_ref / t_ours, frac_elements_differing=diff,
                   max_abs_diff=(y_ours - y_ref).abs().max().item())
    rows.append(row)
    print(json.dumps(row), flush=True)
if args.out:
    json.dump({"peak_GBs": peak, "rows": rows, "note": "fp32 NCHW, mask on, graph replays over rotating buffers > L2; "
               "ours = launch_spatial_quantization of libmcaq_b200.so (vector K3), ref = the reference's "
               "SpatialAdaptiveQuantizationKernel compiled for sm_100"}, open(args.out, "w"), indent=1)
