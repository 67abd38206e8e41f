#!/usr/bin/env python
"""Where does the pipelined step go?  Runs NSLOT x 3 scales concurrently (one stream each) with
(a) only K2, (b) only K1 + K3, (c) the whole chain, and prints the time per step-equivalent."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from golden_util import weights  # noqa: E402
from mcaq_yolo_b200 import constants as K, fused, modules as M, ops  # noqa: E402

NSLOT = int(sys.argv[1]) if len(sys.argv) > 1 else 4
B = 64
dev = torch.device("cuda")
W = weights()
a, m, q = M.build_fixture_modules(W, device=dev)
cm, mp, sm = K.pack_complexity_mlp(a.complexity_mlp), K.pack_mapping_steps(m.mapping_network, 1.0, m.min_bits, m.max_bits), K.pack_soft_mask(q.soft_mask)
shapes = [(64, 80, 80), (128, 40, 40), (256, 20, 20)]
g = torch.Generator(device=dev)
g.manual_seed(1)
data = []
for s in range(NSLOT):
    for C, H, Wd in shapes:
        coarse = torch.randn(B, C, H // 8 + 2, Wd // 8 + 2, device=dev, generator=g)
        x = (torch.nn.functional.interpolate(coarse, size=(H, Wd), mode="bilinear") * 1.6
             + 0.1 * torch.randn(B, C, H, Wd, device=dev, generator=g) + 0.3).to(torch.bfloat16).contiguous()
        ws = fused.ScaleWorkspace(C, dev)
        sp, ap, _ = ops.reduce_planes(x, want_ranges=False)
        ops._call("mcaq_reduce_planes", x.data_ptr(), ops._dtype_code(x), B, C, H, Wd, sp.data_ptr(), ap.data_ptr(),
                  ws.keys.data_ptr(), ops._stream())
        r = ops.morph_fused(sp, ap, C, 8, cm, mp, sm, 1.0, keys=ws.keys)
        y = torch.empty_like(x)
        data.append(dict(x=x, C=C, H=H, W=Wd, ws=ws, sp=sp, ap=ap, r=r, y=y))
torch.cuda.synchronize()
streams = [torch.cuda.Stream() for _ in data]


def k1(d):
    ops._call("mcaq_reduce_planes", d["x"].data_ptr(), ops._dtype_code(d["x"]), B, d["C"], d["H"], d["W"],
              d["sp"].data_ptr(), d["ap"].data_ptr(), d["ws"].keys.data_ptr(), ops._stream())


def k2(d):
    ops.morph_fused(d["sp"], d["ap"], d["C"], 8, cm, mp, sm, 1.0, keys=d["ws"].keys)


def k3(d):
    ops.tile_quantize_ranges(d["x"], d["r"]["bit_map"], d["r"]["packed"], None, None, d["r"]["mask"], out=d["y"])


def measure(name, fns, reps=20):
    main = torch.cuda.current_stream()
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr):
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream())
        joins = []
        for d, st in zip(data, streams):
            st.wait_event(ev)
            with torch.cuda.stream(st):
                for _ in range(4):                       # 4 back-to-back rounds per stream
                    for f in fns:
                        f(d)
                e = torch.cuda.Event()
                e.record(st)
                joins.append(e)
        for e in joins:
            torch.cuda.current_stream().wait_event(e)
    gr.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        gr.replay()
    e1.record()
    torch.cuda.synchronize()
    per_step = e0.elapsed_time(e1) * 1e3 / (reps * 4 * NSLOT)
    print(f"{name:28s} {per_step:7.1f} us per step-equivalent ({NSLOT} slots x 3 scales concurrently)", flush=True)


measure("K2 only", [k2])
measure("K1 + K3 only", [k1, k3])
measure("K1 only", [k1])
measure("K3 only", [k3])
measure("K1 + K2 + K3", [k1, k2, k3])
