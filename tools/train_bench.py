#!/usr/bin/env python
"""Training forms of the hot path (BASELINE configs[3]: YOLOv8n MCAQ training step, 16 images per
GPU): per-kernel time of the fractional-bit forward (2*s bytes / element) and the STE backward
(3*s: read g, read x, write dx), vector vs scalar kernels, with and without the folded feature
distillation term, from CUDA-graph replays over rotating buffers larger than L2; then the whole
module-level training step of the three hooks (analyzer + mapper + quantiser forward, loss, backward
through torch autograd for the three tiny networks) timed eagerly with CUDA events.

    python tools/train_bench.py [--dtype bf16|f32] [--batch 16] [--iters 30] [--out FILE]
Prints one JSON line."""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
from mcaq_yolo_b200 import _lib, ops  # noqa: E402
from mcaq_yolo_b200 import modules as M  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--dtype", default="bf16")
ap.add_argument("--batch", type=int, default=16)
ap.add_argument("--iters", type=int, default=30)
ap.add_argument("--shapes", default="64x80x80,128x40x40,256x20x20")
ap.add_argument("--step-iters", type=int, default=20)
ap.add_argument("--out", default="")
a = ap.parse_args()
dt = torch.bfloat16 if a.dtype == "bf16" else torch.float32
es = 2 if a.dtype == "bf16" else 4
dev = "cuda"
B = a.batch
PEAK = 6541.8
try:
    PEAK = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass
lib = _lib.load()


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
shapes = [tuple(int(v) for v in s.split("x")) for s in a.shapes.split(",")]
for C, H, W in shapes:
    n = B * C * H * W
    nbytes = n * es
    nbuf = max(3, int(500e6 // (3 * nbytes)) + 1)
    xs = [(torch.randn(B, C, H, W, device=dev) * 2 + 0.3).to(dt) for _ in range(nbuf)]
    gs = [torch.randn(B, C, H, W, device=dev).to(dt) for _ in range(nbuf)]
    ts_ = [torch.randn(B, C, H, W, device=dev) for _ in range(nbuf)]
    tile = ops.tile_size(H, 8)
    Ht, Wt = H // tile, W // tile
    bm = torch.rand(B, Ht, Wt, device=dev) * 6 + 2
    m = torch.rand(B, H, W, device=dev) * 0.2 + 0.8
    _, _, keys = ops.reduce_planes(xs[0])
    qt = ops.build_qtable(ops.ranges_decode(keys))
    coef = torch.tensor([1e-6], device=dev)
    row = {"shape": [B, C, H, W], "dtype": a.dtype}

    def rec(name, us, alg):
        row[name] = {"us": round(us, 2), "GBps": round(alg / us / 1e3, 1), "frac": round(alg / us / 1e3 / PEAK, 3)}

    for tag, scalar in (("vec", 0), ("scalar", 1)):
        lib.mcaq_debug_train_scalar(scalar)
        rec(f"fwd_{tag}", timeit(lambda i: ops.tile_quantize_train_fwd(xs[i], bm, qt, m), nbuf, a.iters), 2 * nbytes)
        rec(f"bwd_{tag}", timeit(lambda i: ops.tile_quantize_train_bwd(gs[i], xs[i], bm, qt, m), nbuf, a.iters), 3 * nbytes)
    lib.mcaq_debug_train_scalar(0)
    rec("fwd_kd", timeit(lambda i: ops.tile_quantize_train_fwd_kd(xs[i], bm, qt, m, ts_[i]), nbuf, a.iters),
        2 * nbytes + 4 * n)
    rec("bwd_kd", timeit(lambda i: ops.tile_quantize_train_bwd_kd(gs[i], xs[i], bm, qt, m, ts_[i], coef), nbuf, a.iters),
        3 * nbytes + 4 * n)

    # the same loss composed: forward + torch mse (reads y and teacher again), backward of both
    def composed_fwd(i):
        y = ops.tile_quantize_train_fwd(xs[i], bm, qt, m)
        return torch.nn.functional.mse_loss(y.float(), ts_[i])
    row["fwd_plus_torch_mse_us"] = round(timeit(composed_fwd, nbuf, a.iters), 2)
    rows.append(row)
    del xs, gs, ts_

# ---- module-level training step of the three hooks (eager) --------------------------------------
from golden_util import weights  # noqa: E402

Wt_ = weights()
analyzer, mapper, _ = M.build_fixture_modules(Wt_, device=dev)
analyzer.train(); mapper.train()
quants = []
for _ in shapes:
    _, _, q = M.build_fixture_modules(Wt_, device=dev)
    quants.append(q.train())
params = [p for mod in [analyzer, mapper] + quants for p in mod.parameters()]
NSETS = 4
feats = [[(torch.randn(B, C, H, W, device=dev) * 2 + 0.3).to(dt) for C, H, W in shapes] for _ in range(NSETS)]
teach = [[torch.randn(B, C, H, W, device=dev) for C, H, W in shapes] for _ in range(NSETS)]
gouts = [[torch.randn(B, C, H, W, device=dev).to(dt) * 1e-3 for C, H, W in shapes] for _ in range(NSETS)]


SCALE_STREAMS = [torch.cuda.Stream() for _ in shapes]


def train_step(i, fused_kd=True, scale_streams=False):
    """One step.  `scale_streams`: each scale's hook (and therefore, through autograd's stream affinity, its
    backward) runs on its own stream, like the three scale streams of the inference path."""
    from mcaq_yolo_b200 import train_nets as TN
    k = i % NSETS
    for p in params:
        p.grad = None
    main = torch.cuda.current_stream()
    if scale_streams:
        TN.prepare_step(analyzer.complexity_mlp, mapper.mapping_network)    # shared nodes live on the main stream
    partial, bits_all = [], []
    for si, (x0, t, go, q) in enumerate(zip(feats[k], teach[k], gouts[k], quants)):
        st = SCALE_STREAMS[si] if scale_streams else main
        if scale_streams:
            st.wait_stream(main)
        with torch.cuda.stream(st):
            x = x0.detach().requires_grad_(True)
            q.kd_teacher = t if fused_kd else None
            r = M.mcaq_hook_forward(x, analyzer, mapper, q, temperature=1.0, training=True)
            y = r["features_q"]
            kd = r.get("kd_feature_loss")
            if kd is None:
                kd = torch.nn.functional.mse_loss(y.float(), t)
            part = (y * go).sum().float() + kd / len(shapes)      # stand-in for the detection loss + KD
        if scale_streams:
            part.record_stream(main)
            r["bit_map"].record_stream(main)
        partial.append(part)
        bits_all.append(r["bit_map"])
    if scale_streams:
        for st in SCALE_STREAMS:
            main.wait_stream(st)
    _, lbit, _ = TN.bit_map_losses(bits_all, 4.0)                            # Lbit (models/mcaq_yolo.py:113-118)
    loss = torch.stack(partial).sum() + 0.1 * lbit
    loss.backward()
    return loss


def time_step(fused_kd):
    for i in range(5):
        train_step(i, fused_kd)
    torch.cuda.synchronize()
    ops.LAUNCHES = 0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(a.step_iters):
        train_step(i, fused_kd)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / a.step_iters, ops.LAUNCHES // a.step_iters


def time_step_graph(fused_kd, scale_streams=False):
    """Whole training step (forward, loss, backward) of one input set captured in a CUDA graph per
    set and replayed: removes the host launch cost of the small autograd kernels."""
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for i in range(3):
            train_step(i, fused_kd, scale_streams)
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    graphs = []
    for k in range(NSETS):
        for p in params:
            p.grad = None
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            train_step(k, fused_kd, scale_streams)
        graphs.append(g)
    for g in graphs:
        g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(a.step_iters):
        graphs[i % NSETS].replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / a.step_iters


# the captured step runs first: autograd's AccumulateGrad nodes must not have been created on the
# default stream by an earlier eager step (their stream is replayed inside the capture)
graph_ms, graph_err = None, None
try:
    graph_ms = time_step_graph(True)
except Exception as e:          # noqa: BLE001 -- report, the eager number stands
    graph_err = f"{type(e).__name__}: {str(e)[:200]}"
    torch.cuda.synchronize()
graph_ss_ms, graph_ss_err = None, None
try:
    graph_ss_ms = time_step_graph(True, scale_streams=True)
except Exception as e:          # noqa: BLE001
    graph_ss_err = f"{type(e).__name__}: {str(e)[:200]}"
    torch.cuda.synchronize()
ms_f, launches = time_step(True)
ms_u, _ = time_step(False)
elems = sum(C * H * W for C, H, W in shapes)
out = {"what": "MCAQ training hot path (K1 + phi + nets + fractional quantise fwd/bwd + feature KD), eager, 1 GPU",
       "batch": B, "dtype": a.dtype, "peak_GBps": PEAK, "kernels": rows,
       "train_step_ms_fused_kd": round(ms_f, 3), "train_step_ms_composed_kd": round(ms_u, 3),
       "images_per_s_fused_kd": round(B / ms_f * 1e3, 1),
       "train_step_ms_cuda_graph": None if graph_ms is None else round(graph_ms, 3),
       "images_per_s_cuda_graph": None if graph_ms is None else round(B / graph_ms * 1e3, 1),
       "cuda_graph_error": graph_err,
       "train_step_ms_cuda_graph_scale_streams": None if graph_ss_ms is None else round(graph_ss_ms, 3),
       "images_per_s_cuda_graph_scale_streams": None if graph_ss_ms is None else round(B / graph_ss_ms * 1e3, 1),
       "cuda_graph_scale_streams_error": graph_ss_err,
       "native_launches_per_step": launches,
       "algorithmic_bytes_per_step": 6 * es * elems * B,
       "note": "step = 3 hooks forward (train mode: EMA ranges, continuous bits, soft mask) + loss + backward; the three "
               "tile-level networks, their batch statistics and the bit-map losses are native kernels behind autograd "
               "Functions (train_nets.py); *_scale_streams: one stream per scale, forward and backward"}
line = json.dumps(out)
print(line)
if a.out:
    with open(a.out, "w") as f:
        f.write(line + "\n")
