#!/usr/bin/env python
"""BASELINE configs[3] at N GPUs: MCAQ training step of the three hooks, 16 images per rank (batch 128 over 8 x B200),
bf16 feature maps, with the exchanges that make the sharded step equal the unsharded one:

  * mapper BatchNorm batch statistics of the whole batch, merged INSIDE the mapper kernels over NVLink peer memory
    (no collective launch; csrc/train_nets.cu),
  * EMA per-channel ranges over the whole batch (one NCCL MIN all-reduce of [min, -max] per scale),
  * avg_bits / Lbit / Lsmooth over the global batch (one all-reduce of 6 floats, straight-through local gradient),
  * one flat NCCL all-reduce of the ~32 KB of small-network gradients.

Launch:  torchrun --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/train_dist.py [--steps K]
Prints one JSON line on rank 0 (step ms = max over ranks, CUDA events) and checks, outside the timed region, that
every rank holds identical BatchNorm running statistics and identical all-reduced gradients."""
import argparse
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)

ap = argparse.ArgumentParser()
ap.add_argument("--steps", type=int, default=30)
ap.add_argument("--warmup", type=int, default=5)
ap.add_argument("--batch", type=int, default=16)
ap.add_argument("--out", default="")
ap.add_argument("--graph", action="store_true", help="also time the step as one captured CUDA graph per input set (NCCL "
                                                     "all-reduces and the in-kernel peer exchange captured with it)")
ap.add_argument("--scale-streams", action="store_true", help="with --graph: one stream per scale, forward and backward")
a = ap.parse_args()

world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
out_fd = os.dup(1)
if world > 1:
    os.dup2(2, 1)
    dist.init_process_group("nccl", device_id=dev)

from mcaq_yolo_b200 import modules as M  # noqa: E402
from mcaq_yolo_b200 import train_nets as TN  # noqa: E402
from mcaq_yolo_b200.peer import RangeExchange  # noqa: E402
from golden_util import weights  # noqa: E402

# everything runs on a non-default stream: autograd binds a parameter's AccumulateGrad node to the stream of its first
# backward, and a node bound to the legacy default stream cannot take part in a later graph capture
torch.cuda.set_stream(torch.cuda.Stream(device=dev))
B = a.batch
shapes = [(64, 80, 80), (128, 40, 40), (256, 20, 20)]
W = weights()
analyzer, mapper, _ = M.build_fixture_modules(W, device=dev)
analyzer.train(); mapper.train()
quants = [M.build_fixture_modules(W, device=dev)[2].train() for _ in shapes]
for q in quants:
    q.sync_ranges = world > 1
if world > 1:
    mapper.stat_exchange = RangeExchange.create(128)
nets = [analyzer.complexity_mlp, mapper.mapping_network] + [q.soft_mask.net for q in quants]
params = [p for n in nets for p in n.parameters()]
gen = torch.Generator(device=dev); gen.manual_seed(100 + rank)
NSETS = 4
feats = [[(torch.randn(B, C, H, Wd, device=dev, generator=gen) * 2 + 0.3).bfloat16() for C, H, Wd in shapes] for _ in range(NSETS)]
teach = [[torch.randn(B, C, H, Wd, device=dev, generator=gen) for C, H, Wd in shapes] for _ in range(NSETS)]
gouts = [[(torch.randn(B, C, H, Wd, device=dev, generator=gen) * 1e-3).bfloat16() for C, H, Wd in shapes] for _ in range(NSETS)]


SS = [torch.cuda.Stream(device=dev) for _ in shapes]


def step(i, scale_streams=False):
    k = i % NSETS
    for p in params:
        p.grad = None
    main = torch.cuda.current_stream()
    if scale_streams:
        TN.prepare_step(analyzer.complexity_mlp, mapper.mapping_network)
    parts, bits = [], []
    for si, (x0, t, go, q) in enumerate(zip(feats[k], teach[k], gouts[k], quants)):
        st = SS[si] if scale_streams else main
        if scale_streams:
            st.wait_stream(main)
        with torch.cuda.stream(st):
            x = x0.detach().requires_grad_(True)
            q.kd_teacher = t
            r = M.mcaq_hook_forward(x, analyzer, mapper, q, temperature=1.0, training=True)
            parts.append((r["features_q"] * go).sum().float() + r["kd_feature_loss"] / len(shapes))
        bits.append(r["bit_map"])
    if scale_streams:
        for st in SS:
            main.wait_stream(st)
    avg, lbit, lsm = TN.bit_map_losses(bits, 4.0)           # global batch when world > 1
    loss = torch.stack(parts).sum() + 0.01 * lbit + 0.1 * lsm
    loss.backward()
    TN.allreduce_grads(nets)                                  # one flat ~32 KB all-reduce
    return loss, avg


for i in range(a.warmup):
    step(i)
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(a.steps):
    _, avg = step(i)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / a.steps
t = torch.tensor([ms], device=dev, dtype=torch.float64)
if world > 1:
    dist.all_reduce(t, op=dist.ReduceOp.MAX)

graph_ms, graph_err = None, None
if a.graph:
    try:
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for i in range(3):
                step(i, a.scale_streams)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        graphs = []
        for k in range(NSETS):
            for p in params:
                p.grad = None
            g = torch.cuda.CUDAGraph()
            # thread_local: the NCCL watchdog thread makes CUDA calls of its own while this thread captures
            with torch.cuda.graph(g, capture_error_mode="thread_local"):
                step(k, a.scale_streams)
            graphs.append(g)
        for g in graphs:
            g.replay()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g0.record()
        for i in range(a.steps):
            graphs[i % NSETS].replay()
        g1.record()
        torch.cuda.synchronize()
        tg = torch.tensor([g0.elapsed_time(g1) / a.steps], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(tg, op=dist.ReduceOp.MAX)
        graph_ms = float(tg.item())
    except Exception as e:      # noqa: BLE001 -- the eager number stands
        graph_err = f"{type(e).__name__}: {str(e)[:200]}"
        torch.cuda.synchronize()

if a.graph:                      # the consistency check below looks at the gradients of a plain eager step
    torch.cuda.synchronize()
    if graph_err is not None:
        sys.stderr.write("[train_dist] capture failed: %s\n" % graph_err)
    _, avg = step(0)
    torch.cuda.synchronize()

# ---- consistency over ranks, outside the timed region ----------------------------------------------------------
ok = 1
if world > 1:
    mapper.stat_exchange.check()
    for bn in (mapper.mapping_network[1], mapper.mapping_network[4], mapper.mapping_network[7]):
        for buf in (bn.running_mean, bn.running_var):
            lo, hi = buf.clone(), buf.clone()
            dist.all_reduce(lo, op=dist.ReduceOp.MIN); dist.all_reduce(hi, op=dist.ReduceOp.MAX)
            ok &= int(torch.equal(lo, hi))                 # bit-identical on every rank: same merged statistics
    for p in params:
        lo, hi = p.grad.clone(), p.grad.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN); dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        ok &= int(torch.equal(lo, hi))
    for q in quants:
        lo, hi = q.running_min.clone(), q.running_min.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN); dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        ok &= int(torch.equal(lo, hi))
    flag = torch.tensor([ok], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    ok = int(flag.item())
line = {"what": "MCAQ training step of the three hooks (configs[3] share: %d images per GPU, bf16), eager launches" % B,
        "n_gpus": world, "global_batch": B * world, "ms_per_step": float(t.item()),
        "images_per_s": B * world / float(t.item()) * 1e3, "avg_bits": float(avg),
        "exchanges": "SyncBN statistics inside the mapper kernels over peer memory; ranges MIN all-reduce; avg_bits/TV "
                     "all-reduce (6 floats); one flat gradient all-reduce" if world > 1 else "none (single GPU)",
        "rank_consistency": "ok" if ok else "FAILED"}
if a.graph:
    line.update({"captured_ms_per_step": graph_ms, "captured_images_per_s": None if graph_ms is None else B * world / graph_ms * 1e3,
                 "captured_streams": "one per scale" if a.scale_streams else "one", "capture_error": graph_err})
if rank == 0:
    os.write(out_fd, (json.dumps(line) + "\n").encode())
    if a.out:
        open(a.out, "w").write(json.dumps(line) + "\n")
if world > 1:
    import threading, time
    threading.Thread(target=lambda: (time.sleep(20), os._exit(0)), daemon=True).start()
    dist.barrier()
    dist.destroy_process_group()
