#!/usr/bin/env python
"""Host <-> device copy bandwidth per rank with ALL ranks copying at once (pinned buffers, 92 MB H2D + 92 MB D2H per
iteration on two streams, no kernels): the ceiling of bench.py's `e2e` number at N GPUs.
  torchrun --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/pcie_probe.py"""
import json
import os
import time

import torch
import torch.distributed as dist

world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
nbytes = 91_750_400
h_in = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
h_out = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
d_in = torch.empty(nbytes, dtype=torch.uint8, device=dev)
d_out = torch.empty(nbytes, dtype=torch.uint8, device=dev)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def run(n, both=True, up=True):
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(n):
        if both or up:
            with torch.cuda.stream(s1):
                d_in.copy_(h_in, non_blocking=True)
        if both or not up:
            with torch.cuda.stream(s2):
                h_out.copy_(d_out, non_blocking=True)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([dt], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt = float(t.item())
    return n * nbytes / dt / 1e9


run(5)
res = {"n_gpus": world, "h2d_only_GBs_per_gpu": run(20, both=False, up=True), "d2h_only_GBs_per_gpu": run(20, both=False, up=False),
       "duplex_GBs_per_gpu_each_way": run(20)}
res["images_per_s_ceiling_per_gpu"] = res["duplex_GBs_per_gpu_each_way"] * 1e9 / (nbytes / 64)
if rank == 0:
    print(json.dumps(res), flush=True)
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
