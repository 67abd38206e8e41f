#!/usr/bin/env python
"""K2 (morph_fused) timing per shape and cluster split, CUDA-graph replays, CUDA events.
Usage: python tools/k2_bench.py [batch] [splits e.g. 1,2,4]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from mcaq_yolo_b200 import ops, _lib, constants as K, modules as M  # noqa: E402
from golden_util import weights  # noqa: E402

lib = _lib.load()
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
splits = [int(v) for v in sys.argv[2].split(",")] if len(sys.argv) > 2 else [1, 2, 4]
a, m, q = M.build_fixture_modules(weights(), "cuda")
cm, mp, sm = K.pack_complexity_mlp(a.complexity_mlp), K.pack_mapping_steps(m.mapping_network, 1.0, m.min_bits, m.max_bits), K.pack_soft_mask(q.soft_mask)
for (C, H) in ((64, 80), (128, 40), (256, 20), (128, 160), (256, 80), (512, 40)):
    Bh = B if H < 160 else max(1, B // 2)
    x = torch.nn.functional.interpolate(torch.randn(Bh, C, H // 8, H // 8, device="cuda"), size=(H, H), mode="bicubic")
    x = (x + 0.1 * torch.randn_like(x)).to(torch.bfloat16)
    s, ab, k = ops.reduce_planes(x)
    for ns in splits:
        lib.mcaq_debug_cluster_split(ns)
        try:
            for _ in range(3):
                ops.morph_fused(s, ab, C, 8, cm, mp, sm, 1.0)
        except RuntimeError as e:
            print(f"K2 C={C:3d} H={H:3d} split={ns}: {e}")
            continue
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for _ in range(10):
                r = ops.morph_fused(s, ab, C, 8, cm, mp, sm, 1.0)
        ts = []
        for _ in range(10):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) * 100)
        ts.sort()
        print(f"K2 C={C:3d} H={H:3d} B={Bh:3d} split={ns}: median {ts[5]:7.1f} us  min {ts[0]:7.1f} us", flush=True)
    lib.mcaq_debug_cluster_split(0)
