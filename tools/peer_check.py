#!/usr/bin/env python
"""Multi-GPU check of the in-kernel range exchange (run under torchrun, one rank per GPU):
every rank holds the same full batch; each quantizes its shard through FusedHotPath with a
RangeExchange and the gathered result must equal the unsharded single-GPU result bit for bit."""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from golden_util import weights  # noqa: E402
from mcaq_yolo_b200 import modules as M  # noqa: E402
from mcaq_yolo_b200.fused import FusedHotPath  # noqa: E402
from mcaq_yolo_b200.peer import RangeExchange  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
W = weights()
a, m, _ = M.build_fixture_modules(W, device=dev)
shapes = [(64, 80, 80), (128, 40, 40), (256, 20, 20)]
qs = [M.build_fixture_modules(W, device=dev)[2] for _ in shapes]
B = 4 * world
g = torch.Generator(device=dev)
g.manual_seed(7)                                  # same data on every rank
ok = True
ex = [RangeExchange.create(C) for C, _, _ in shapes]
sharded = FusedHotPath(a, m, qs, exchanges=ex)
full = FusedHotPath(a, m, qs)
for step in range(4):
    feats = [(torch.randn(B, C, H, Wd, device=dev, generator=g) * (1 + step) + 0.1 * step).to(torch.bfloat16)
             for C, H, Wd in shapes]
    for f in feats:                               # rank-dependent ranges
        for r in range(world):
            f[r * 4:(r + 1) * 4] *= (1.0 + 0.3 * r)
    ref = full.run(feats)
    mine = sharded.run([f[rank * 4:(rank + 1) * 4].contiguous() for f in feats])
    torch.cuda.synchronize()
    for i, (r_, m_) in enumerate(zip(ref, mine)):
        same = torch.equal(r_["features_q"][rank * 4:(rank + 1) * 4], m_["features_q"]) and \
            torch.equal(r_["bit_map"][rank * 4:(rank + 1) * 4], m_["bit_map"])
        ok = ok and same
        if not same:
            print(f"rank {rank} step {step} scale {i}: MISMATCH", flush=True)
t = torch.tensor([1 if ok else 0], device=dev)
dist.all_reduce(t, op=dist.ReduceOp.MIN)
if rank == 0:
    print("peer exchange check:", "OK (sharded == unsharded, bit exact)" if int(t.item()) else "FAILED", flush=True)
dist.barrier()
dist.destroy_process_group()
