#!/usr/bin/env python
"""bench.py -- MCAQ complexity + quantize hot path on B200.

    python bench.py --gpus N --steps K --warmup W            (N > 1: launched under torchrun)
    python bench.py --impl reference ...                     (CPU arm: oracle port on host cores)

A step = one pass of the hot path (K1 channel sweep, K2 morphology, complexity MLP, bit mapper,
soft mask, K3 quantize) over the C3/C4/C5 feature maps of one batch.  Workload at every N is
BASELINE.json configs[1] per GPU: YOLOv8n @ 640x640, batch 64, bf16 feature maps
(64x80x80, 128x40x40, 256x20x20), grid 8, bits 2..8, synthetic data, fixture weights.
Prints ONE JSON line (rank 0).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

METRIC = "MCAQ complexity+quantize throughput (C3/C4/C5 hooks, whole job)"
UNIT = "images/s"

WORKLOADS = {
    # name: (per-GPU batch, [(C, H, W)...], dtype, grid)
    "yolov8n_640_b64_bf16": (64, [(64, 80, 80), (128, 40, 40), (256, 20, 20)], "bf16", 8),
    "yolov8s_1280_b32_f32": (32, [(128, 160, 160), (256, 80, 80), (512, 40, 40)], "f32", 8),
    "yolov8n_640_b1_f32": (1, [(64, 80, 80), (128, 40, 40), (256, 20, 20)], "f32", 8),
}
# model-level workloads (north_star reporting): synthetic images through a random-init minimal YOLOv8 driven by the
# UNMODIFIED reference MCAQYOLO (tests/harness + baseline/_ref), native hooks installed; name: (scale, size, batch, dtype)
MODEL_WORKLOADS = {
    "model_v8n_640_b64": ("n", 640, 64, "bf16"),
    "model_v8s_1280_b32": ("s", 1280, 32, "f32"),
    "model_v8n_640_b1": ("n", 640, 1, "f32"),
}
INPUT_SETS = 4          # rotating input sets: 4 x 92 MB > 126 MB L2, so no step starts L2-warm


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--workload", default="yolov8n_640_b64_bf16", choices=sorted(WORKLOADS) + sorted(MODEL_WORKLOADS))
    ap.add_argument("--no-graph", action="store_true", help="time eager launches instead of CUDA-graph replays")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-streams", action="store_true", help="run the three scales on one stream")
    ap.add_argument("--unfused", action="store_true", help="module-by-module path (9 launches per scale)")
    ap.add_argument("--nccl-in-graph", action="store_true", help="capture the range all-reduce inside one graph per step")
    ap.add_argument("--nccl-ranges", action="store_true",
                    help="multi-GPU: merge the ranges with an NCCL all-reduce between K2 and K3 instead of the "
                         "in-kernel peer-memory exchange")
    ap.add_argument("--sharded", action="store_true", help="use the multi-rank phase split even at world size 1")
    ap.add_argument("--inflight", type=int, default=4, choices=[1, 2, 4, 8],
                    help="steps in flight: consecutive steps alternate between this many streams (each with its own "
                         "workspace), so the latency-bound morphology kernel of one step overlaps the HBM sweeps of "
                         "its neighbours; 1 = strictly serial steps")
    ap.add_argument("--input-sets", type=int, default=0, help="rotating input sets (default 4, or --inflight if larger)")
    ap.add_argument("--frozen-ranges", action="store_true",
                    help="calibrate once on the first input set, then freeze the per-channel ranges (the paper's "
                         "deployment mode, quantization.py:647-649): K1 skips the range reduction and no exchange "
                         "between ranks is needed; the default is the reference's un-calibrated eval mode "
                         "(ranges of the current batch)")
    ap.add_argument("--no-k2-priority", dest="k2_priority", action="store_false",
                    help="keep the morphology kernel on its scale's stream (default: a high-priority stream per scale, "
                         "fork / join by events: its few long-lived CTAs are placed ahead of the bandwidth kernels' -- "
                         "measured 0.0906 -> 0.0833 ms/step)")
    ap.add_argument("--no-secondary", action="store_true",
                    help="skip the short run of the per-GPU share of BASELINE configs[2] (YOLOv8s @ 1280, 32 images, fp32) "
                         "that the default single-GPU run appends as `secondary`")
    ap.add_argument("--timeline", default="", help="write the CUPTI kernel timeline (start us, duration us, stream, kernel) of 16 "
                                                   "untimed steps of the same loop to this file (tuning aid)")
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--ref-budget", type=float, default=120.0, help="--impl reference: seconds of host work for the whole run")
    return ap.parse_args()


# ----------------------------------------------------------------------------- CPU arm (oracle)
def _oracle_worker(job):
    """One bounded sample: the oracle's hook_forward over the three scales for `nimg` images."""
    shapes, nimg, seed, grid = job
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import numpy as np
    import mcaq_oracle as o
    from golden_util import weights
    from inputs import feature_map
    W = weights()
    xs = [feature_map("smooth", nimg, C, H, Wd, seed + i) for i, (C, H, Wd) in enumerate(shapes)]
    t0 = time.perf_counter()
    for x in xs:
        o.hook_forward(x, W["analyzer"], W["mapper"], W["quantizer"], grid, 1.0)
    return time.perf_counter() - t0


_REF_STATE = {}


def _reference_available():
    """True when an importable copy of the UNMODIFIED reference is reachable: /root/reference in the build
    container, or baseline/_ref (pip --target install, git-ignored, travels to the GPU box with the snapshot)."""
    from harness import ref_model
    return ref_model.reference_root() is not None


def _reference_worker(job):
    """One bounded sample through the reference's OWN modules on CPU (core/morphology.py, bit_allocation.py,
    quantization.py: the pure-PyTorch path, one torch thread per worker process): analyzer -> mapper ->
    quantizer(training=False) over the three scales for `nimg` images."""
    shapes, nimg, seed, grid = job
    import torch
    if not _REF_STATE:
        torch.set_num_threads(1)
        from harness import ref_model
        pkg = ref_model.load(with_model=False)
        from mcaq_yolo.core import bit_allocation, morphology, quantization
        from golden_util import weights
        W = weights()
        sd = lambda d: {k: torch.as_tensor(v) for k, v in d.items()}      # noqa: E731
        A = morphology.MorphologicalComplexityAnalyzer(grid_size=grid, device="cpu")
        A.load_state_dict(sd(W["analyzer"]))
        Mp = bit_allocation.ComplexityToBitMappingNetwork()
        Mp.load_state_dict(sd(W["mapper"]))
        Qs = []
        for _ in shapes:
            Q = quantization.SpatialAdaptiveQuantization(calibration_mode="minmax", smooth_transitions=True, per_channel=True)
            Q.load_state_dict(sd(W["quantizer"]))
            Qs.append(Q.eval())
        _REF_STATE.update(A=A.eval(), M=Mp.eval(), Q=Qs, pkg=pkg)
    from inputs import feature_map
    xs = [torch.from_numpy(feature_map("smooth", nimg, C, H, Wd, seed + i)) for i, (C, H, Wd) in enumerate(shapes)]
    t0 = time.perf_counter()
    with torch.no_grad():
        for x, Q in zip(xs, _REF_STATE["Q"]):
            c = _REF_STATE["A"](x)
            bm = _REF_STATE["M"](c, 1.0)
            Q(x, bm, training=False)
    return time.perf_counter() - t0


def cpu_sample(shapes, grid, seconds, procs):
    """images/s of the oracle port on `procs` host processes (each its own 2-image batches)."""
    import multiprocessing as mp
    nimg = 2
    t_one = _oracle_worker((shapes, nimg, 0, grid))
    rounds = max(1, int(seconds / max(t_one, 1e-3)))
    if procs == 1:
        t0 = time.perf_counter()
        for r in range(rounds):
            _oracle_worker((shapes, nimg, r, grid))
        dt = time.perf_counter() - t0
        return nimg * rounds / dt, nimg * rounds, dt
    ctx = mp.get_context("spawn")
    with ctx.Pool(procs) as pool:
        pool.map(_oracle_worker, [(shapes, nimg, 1000 + i, grid) for i in range(procs)])   # warm imports
        t0 = time.perf_counter()
        pool.map(_oracle_worker, [(shapes, nimg, i, grid) for i in range(procs * rounds)])
        dt = time.perf_counter() - t0
    return nimg * procs * rounds / dt, nimg * procs * rounds, dt


def run_reference(args):
    """The reference's CPU implementation of the path on all host cores: the reference's OWN torch modules when a
    copy is reachable (baseline/_ref travels to the GPU box), else the numpy oracle port; each step is a bounded
    sample and the whole run is sized to about two minutes whatever --steps / --warmup are."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import multiprocessing as mp
    B, shapes, dtype, grid = WORKLOADS[args.workload]
    cores = os.cpu_count() or 1
    procs = max(1, min(cores, 64))
    use_ref = _reference_available()
    worker = _reference_worker if use_ref else _oracle_worker
    nimg = 2
    nsteps = max(1, args.steps + args.warmup)
    budget = float(args.ref_budget)
    ctx = mp.get_context("spawn")
    vals = []
    with ctx.Pool(procs) as pool:
        t0 = time.perf_counter()
        pool.map(worker, [(shapes, nimg, 1000 + i, grid) for i in range(procs)])   # imports + one round
        t_round = time.perf_counter() - t0
        t1 = time.perf_counter()
        pool.map(worker, [(shapes, nimg, 2000 + i, grid) for i in range(procs)])
        t_round = min(t_round, time.perf_counter() - t1)
        rounds = max(1, int(budget / nsteps / max(t_round, 1e-3)))
        if rounds * t_round * nsteps > 2.5 * budget:         # even one round per step is too long: fewer jobs
            rounds = 1
        for i in range(nsteps):
            t0 = time.perf_counter()
            pool.map(worker, [(shapes, nimg, 10 * i + j, grid) for j in range(procs * rounds)])
            dt = time.perf_counter() - t0
            if i >= args.warmup:
                vals.append((nimg * procs * rounds, dt))
            if time.perf_counter() - t1 > 3.0 * budget and len(vals) >= 1:
                break                                            # hard stop: never run for many minutes
    nimg_tot = sum(v[0] for v in vals)
    dt = sum(v[1] for v in vals)
    value = nimg_tot / dt
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / max(1, len(vals)),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": args.workload, "grid": grid,
                   "note": ("the UNMODIFIED reference's own modules (core/morphology.py, bit_allocation.py, quantization.py: "
                            "its pure-PyTorch CPU path) from baseline/_ref or /root/reference, one torch thread per process"
                            if use_ref else
                            "oracle port (numpy) of the reference's pure-PyTorch CPU path (no copy of the reference reachable)"),
                   "timed_steps": len(vals)},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": procs, "kind": "reference" if use_ref else "port",
                         "sample": f"{nimg_tot} images in 2-image batches over {procs} processes, {dt:.1f} s"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------- clocks sampler
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap,power.draw")

    def __init__(self, index):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "50", "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
                for n, v in zip(names, r[2:6]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                pass
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}



# ----------------------------------------------------------------------------- host placement (e2e)
def bind_to_gpu_numa(local):
    """Pin this process to the CPUs next to its GPU BEFORE any pinned buffer is allocated (first touch places
    the pages on that NUMA node): with 8 ranks on a two-socket host, pinned buffers on the wrong socket make
    every H2D / D2H copy cross the inter-socket link.  Returns a small evidence dict for the JSON line."""
    info = {"numa_bound": False}
    try:
        import pynvml
        pynvml.nvmlInit()
        idx = local
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        if vis:
            try:
                idx = int(vis.split(",")[local])
            except ValueError:
                idx = local
        h = pynvml.nvmlDeviceGetHandleByIndex(idx)
        ncpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
        cpus = [64 * w + b for w, word in enumerate(words) for b in range(64) if (int(word) >> b) & 1]
        cpus = [c for c in cpus if c < ncpu]
        try:
            info["pcie_gen"] = int(pynvml.nvmlDeviceGetCurrPcieLinkGeneration(h))
            info["pcie_width"] = int(pynvml.nvmlDeviceGetCurrPcieLinkWidth(h))
        except Exception:
            pass
        if cpus:
            os.sched_setaffinity(0, cpus)
            info.update(numa_bound=True, cpus=len(cpus), first_cpu=cpus[0])
        bus = pynvml.nvmlDeviceGetPciInfo(h).busId
        bus = bus.decode() if isinstance(bus, bytes) else bus
        node = os.path.join("/sys/bus/pci/devices", bus.lower()[-12:], "numa_node")
        if os.path.exists(node):
            info["numa_node"] = int(open(node).read().strip())
    except Exception as e:       # no NVML / no permission: run unbound and say so
        info["error"] = f"{type(e).__name__}: {e}"[:120]
    return info

# ----------------------------------------------------------------------------- native arm
def run_native(args):
    global INPUT_SETS
    INPUT_SETS = max(args.input_sets or INPUT_SETS, args.inflight)
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback for the native arm)"
    host_info = bind_to_gpu_numa(local)
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    out_fd = os.dup(1)
    if world > 1:
        os.dup2(2, 1)            # NCCL prints its version banner on stdout: keep stdout for the ONE JSON line
        dist.init_process_group("nccl", device_id=dev)
    from mcaq_yolo_b200 import modules as M
    from mcaq_yolo_b200 import ops
    from golden_util import weights

    B, shapes, dtype_name, grid = WORKLOADS[args.workload]
    tdtype = torch.bfloat16 if dtype_name == "bf16" else torch.float32
    esize = 2 if dtype_name == "bf16" else 4
    W = weights()
    analyzer, mapper, _ = M.build_fixture_modules(W, device=dev, grid_size=grid)
    quantizers = []
    for _ in shapes:
        _, _, q = M.build_fixture_modules(W, device=dev, grid_size=grid)
        quantizers.append(q)

    # synthetic low-frequency feature maps, seeded per rank; INPUT_SETS rotating copies
    gen = torch.Generator(device=dev)
    gen.manual_seed(1234 + rank)

    def synth(C, H, Wd):
        coarse = torch.randn(B, C, H // 8 + 2, Wd // 8 + 2, device=dev, generator=gen)
        up = torch.nn.functional.interpolate(coarse, size=(H, Wd), mode="bilinear", align_corners=False)
        bias = torch.randn(1, C, 1, 1, device=dev, generator=gen) * 0.5
        x = up * 1.6 + 0.1 * torch.randn(B, C, H, Wd, device=dev, generator=gen) + bias + 0.3
        return x.to(tdtype).contiguous()

    sets = [[synth(*s) for s in shapes] for _ in range(INPUT_SETS)]
    if args.frozen_ranges:
        # MCAQYOLO.calibrate (models/mcaq_yolo.py:475-508): EMA statistics from calibration batches, then freeze
        with torch.no_grad():
            for x, q in zip(sets[0], quantizers):
                q.update_running_stats(x)
                q.freeze_calibration()
    elems = sum(C * H * Wd for C, H, Wd in shapes)
    alg_bytes_step = 3 * esize * elems * B            # K1 read + K3 read + K3 write

    from mcaq_yolo_b200.fused import FusedHotPath
    nslot = 1 if (args.unfused or args.nccl_ranges or args.sharded) else args.inflight
    hots = []
    for _slot in range(nslot):
        exchanges = None
        if world > 1 and not args.nccl_ranges and not args.unfused and not args.frozen_ranges:
            # batch sharded over the GPUs of the node: ranges merged inside K2 over peer memory
            from mcaq_yolo_b200.peer import RangeExchange
            try:
                exchanges = [RangeExchange.create(C) for C, _, _ in shapes]
            except RuntimeError as e:        # raised on every rank alike: use the NCCL all-reduce path
                sys.stderr.write(f"[bench] {e}; falling back to --nccl-ranges\n")
                args.nccl_ranges = True
                nslot = 1
                hots = []
                exchanges = None
                hots.append(FusedHotPath(analyzer, mapper, quantizers, temperature=1.0, streams=not args.no_streams))
                break
        hots.append(FusedHotPath(analyzer, mapper, quantizers, temperature=1.0, streams=not args.no_streams,
                                 exchanges=exchanges, latency=(nslot == 1), k2_priority=args.k2_priority))
    hot = hots[0]
    slot_streams = [torch.cuda.Stream(device=dev) for _ in range(nslot)]

    sharded = None
    if ((world > 1 and args.nccl_ranges) or args.sharded) and not args.unfused:
        # multi-rank: the range all-reduce is kept out of the captured graphs (fused.ShardedHotPath)
        from mcaq_yolo_b200.fused import ShardedHotPath
        sharded = ShardedHotPath(analyzer, mapper, quantizers, shapes, dev, 1.0)

    def step(feats, slot=0):
        if args.unfused:
            return [M.mcaq_hook_forward(x, analyzer, mapper, q, temperature=1.0) for x, q in zip(feats, quantizers)]
        if sharded is not None:
            return sharded.run(feats)
        return hots[slot].run(feats)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    with torch.no_grad():
        # ---- warm-up (eager) primes host-side caches, then optional CUDA-graph capture ------
        for i in range(max(3, args.warmup)):
            step(sets[i % INPUT_SETS], (i % INPUT_SETS) % nslot)
        torch.cuda.synchronize()
        ops.LAUNCHES = 0
        step(sets[0])
        launches_per_step = ops.LAUNCHES
        graphs = None
        if sharded is not None and args.nccl_in_graph and not args.no_graph:
            graphs = []
            keep = []
            for s_ in sets:
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    keep.append(sharded.run(s_))
                graphs.append((g,))
        elif sharded is not None:
            seg = []
            if not args.no_graph:
                for s_ in sets:
                    gA, gB1, gB2 = torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph()
                    with torch.cuda.graph(gA):
                        planes = sharded.sweep(s_)
                    with torch.cuda.graph(gB1):
                        nets_out = sharded.nets(s_, planes)
                    with torch.cuda.graph(gB2):
                        recs = sharded.quantize(s_, nets_out)
                    seg.append((gA, gB1, gB2, planes, nets_out, recs))
                graphs = seg
        elif not args.no_graph:
            try:
                graphs, keep = [], []
                for j, s in enumerate(sets):
                    g = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(g, stream=slot_streams[j % nslot]):
                        keep.append(step(s, j % nslot))
                    graphs.append(g)
                for j, g in enumerate(graphs):
                    with torch.cuda.stream(slot_streams[j % nslot]):
                        g.replay()
                torch.cuda.synchronize()
            except Exception as e:       # capture unsupported: time eager launches
                sys.stderr.write(f"[bench] CUDA-graph capture failed ({type(e).__name__}: {e}); timing eager\n")
                graphs = None
                torch.cuda.synchronize()

        def run_step(i):
            if sharded is not None:
                if graphs is None:
                    sharded.run(sets[i % INPUT_SETS])
                    return
                if len(graphs[i % INPUT_SETS]) == 1:
                    graphs[i % INPUT_SETS][0].replay()
                    return
                gA, gB1, gB2 = graphs[i % INPUT_SETS][:3]
                gA.replay()
                done = sharded.exchange()
                gB1.replay()
                torch.cuda.current_stream().wait_event(done)
                gB2.replay()
            else:
                j = i % INPUT_SETS
                with torch.cuda.stream(slot_streams[j % nslot]):     # steps alternate between the slot streams
                    if graphs is not None:
                        graphs[j].replay()
                    else:
                        step(sets[j], j % nslot)

        def fork_slots():
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream())
            for st in slot_streams:
                st.wait_event(ev)

        def join_slots():
            for st in slot_streams:
                ev = torch.cuda.Event()
                ev.record(st)
                torch.cuda.current_stream().wait_event(ev)

        for i in range(args.warmup):
            run_step(i)
        # ---- timed region: exactly K steps, device time, max over ranks -----------------------
        sampler = ClockSampler(local)
        if rank == 0:
            sampler.start()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fork_slots()
        for i in range(args.steps):
            run_step(i)
        join_slots()
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())
        ms_per_step = ms_total / args.steps
        # nvidia-smi samples every 50 ms: when the K timed steps are shorter than that, keep the same
        # load running (untimed) for half a second so the clock / throttle samples are taken under it
        # (the count derives from the all-reduced time: every rank must run the same number of steps)
        burst = 0
        if ms_total < 400.0:
            burst = int(min(20000, 500.0 / max(ms_per_step, 1e-3)))
            burst -= burst % INPUT_SETS
            fork_slots()
            for i in range(burst):
                run_step(i)
            join_slots()
            barrier()
        clocks = sampler.stop() if rank == 0 else None
        if clocks is not None:
            clocks["note"] = ("sampled over the timed region" if burst == 0 else
                              "sampled over the timed region plus %d untimed steps of the same load" % burst)
        value = world * B * args.steps / (ms_total * 1e-3)
        if args.timeline and rank == 0:
            from torch.profiler import ProfilerActivity, profile
            with profile(activities=[ProfilerActivity.CUDA]) as prof:
                fork_slots()
                for i in range(16):
                    run_step(i)
                join_slots()
                torch.cuda.synchronize()
            evs = [e for e in prof.events() if e.device_type.name == "CUDA" and e.time_range.end > e.time_range.start]
            t0_ = min(e.time_range.start for e in evs)
            with open(args.timeline, "w") as f:
                for e in sorted(evs, key=lambda e: e.time_range.start):
                    f.write("%9.1f %7.1f s%s %s\n" % (e.time_range.start - t0_, e.time_range.end - e.time_range.start,
                                                      getattr(e, "device_resource_id", -1), e.name[:70]))

        # ---- roofline attribution: each kernel of each scale timed live with CUDA events around a
        #      CUDA-graph replay of INPUT_SETS back-to-back launches over the rotating inputs (cold
        #      L2; a graph keeps the queue full so host launch latency is not measured)
        from mcaq_yolo_b200 import constants as KC
        from mcaq_yolo_b200.fused import ScaleWorkspace

        ROUNDS = 4          # the rotation is walked ROUNDS times per replay: the events then bracket 4 x
                            # INPUT_SETS launches, so the replay's own start-up is amortised like in a long step

        def graph_time(fn, reps=10):
            for i in range(INPUT_SETS):
                fn(i)
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                for _r in range(ROUNDS):
                    for i in range(INPUT_SETS):
                        fn(i)
            g.replay()
            torch.cuda.synchronize()
            ts = []
            for _ in range(reps):
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                g.replay()
                b.record()
                torch.cuda.synchronize()
                ts.append(a.elapsed_time(b) / (INPUT_SETS * ROUNDS))
            return statistics.mean(ts)

        cm = KC.pack_complexity_mlp(analyzer.complexity_mlp)
        mp_ = KC.pack_mapping_steps(mapper.mapping_network, 1.0, mapper.min_bits, mapper.max_bits)
        kern_ms = {}
        for si, (C, H, Wd) in enumerate(shapes):
            xs_ = [sets[j][si] for j in range(INPUT_SETS)]
            ws = ScaleWorkspace(C, dev)
            sm_ = KC.pack_soft_mask(quantizers[si].soft_mask)
            planes = [ops.reduce_planes(x, want_ranges=False)[:2] for x in xs_]
            fr = [ops.morph_fused(p[0], p[1], C, grid, cm, mp_, sm_, 1.0) for p in planes]
            pk = ops.ranges_decode(ops.reduce_planes(xs_[0])[2])
            ys_ = [torch.empty_like(x) for x in xs_]
            tag = "C%d" % (3 + si)
            kern_ms["K1_reduce_planes_" + tag] = graph_time(
                lambda i: ops._call("mcaq_reduce_planes", xs_[i].data_ptr(), ops._dtype_code(xs_[i]), B, C, H, Wd,
                                    planes[i][0].data_ptr(), planes[i][1].data_ptr(), ws.keys.data_ptr(), ops._stream()))
            kern_ms["K2_morph_fused_" + tag] = graph_time(
                lambda i: ops.morph_fused(planes[i][0], planes[i][1], C, grid, cm, mp_, sm_, 1.0, keys=ws.keys))
            kern_ms["K3_tile_quantize_" + tag] = graph_time(
                lambda i: ops.tile_quantize_ranges(xs_[i], fr[i]["bit_map"], pk, None, None, fr[i]["mask"], out=ys_[i]))
        tot_kernel_ms = sum(kern_ms.values())
        shares = {k: round(v / tot_kernel_ms, 4) for k, v in kern_ms.items()}

        peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.exists(peaks_path):
            peak = float(json.load(open(peaks_path))["hbm_gbs"])
            peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)"
        else:
            peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"

        # ---- the hooks as a model delivers them: C3, C4, C5 serialised on ONE stream by the backbone, one
        #      forward in flight (FusedMcaqHook: latency split policy), as a CUDA-graph replay
        serial = FusedHotPath(analyzer, mapper, quantizers, temperature=1.0, streams=False, latency=True)
        serial_ms = graph_time(lambda i: serial.run(sets[i]))
        from mcaq_yolo_b200 import _lib as _L
        _L.load().mcaq_morph_policy(0 if nslot > 1 else 1)

        # per-kernel algorithmic bandwidth of the six HBM sweeps (K1 reads s*BCHW, K3 reads + writes 2*s*BCHW)
        hbm = {}
        for si, (C, H, Wd) in enumerate(shapes):
            tag = "C%d" % (3 + si)
            nb = esize * B * C * H * Wd
            for nm, mult in (("K1_reduce_planes_", 1), ("K3_tile_quantize_", 2)):
                ms_ = kern_ms[nm + tag]
                hbm[nm + tag] = {"algorithmic_bytes": mult * nb, "avg_launch_ms": round(ms_, 5),
                                 "achieved": mult * nb / (ms_ * 1e-3) / 1e9, "frac": mult * nb / (ms_ * 1e-3) / 1e9 / peak}
        dom = max(kern_ms, key=kern_ms.get)
        # DRAM traffic of the step's HBM kernels from a committed `ncu --set full` capture taken over rotating
        # buffers (steady state); ARCHIVAL -- not measured in this run -- and None when absent
        traffic, traffic_note = None, None
        cap = os.path.join(ROOT, "profiles", "r02_ncu_step_%s_dram.json" % dtype_name)
        if args.workload == "yolov8n_640_b64_bf16" and os.path.exists(cap):
            try:
                tj = json.load(open(cap))
                traffic = float(tj["dram_bytes_per_step"])
                traffic_note = "archival: %s (%s)" % (os.path.basename(cap), tj.get("how", ""))
            except Exception:
                traffic = None
        achieved = alg_bytes_step / (ms_per_step * 1e-3) / 1e9
        roofline = {
            "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
            "traffic": traffic, "traffic_note": traffic_note,
            "kernel": "whole fused path, one step = K1 reduce_planes + K2 morph_fused + K3 tile_quantize_vec on C3/C4/C5 "
                      "(%d launches, %d steps in flight)" % (launches_per_step, nslot),
            "algorithmic_bytes_per_launch": alg_bytes_step, "avg_launch_ms": ms_per_step, "peak_source": peak_src,
            "how": "3*s*sum(CHW)*B bytes per step (SURVEY 8d) / the timed region's ms per step (CUDA events, max over ranks)",
            "whole_step": {"algorithmic_bytes": alg_bytes_step, "achieved": achieved, "frac": achieved / peak},
            "serial_hook": {"ms_per_forward": serial_ms, "achieved": alg_bytes_step / (serial_ms * 1e-3) / 1e9,
                            "frac": alg_bytes_step / (serial_ms * 1e-3) / 1e9 / peak,
                            "how": "the three hooks serialised on one stream, one forward in flight (what "
                                   "FusedMcaqHook delivers inside a backbone), graph replay over rotating inputs"},
            "dominant_kernel": {"name": dom, "time_share": shares[dom], "avg_launch_ms": round(kern_ms[dom], 5),
                                "note": "morph_fused (K2) is on-chip work on the (B,H,W) planes: < 2 % of the step's HBM "
                                        "bytes, bound by instruction issue / latency, overlapped with the sweeps of "
                                        "other scales and steps"},
            "hbm_kernels": hbm,
            "per_kernel_how": "CUDA events around a graph replay of %d launches over %d rotating inputs (cold L2)" % (INPUT_SETS * ROUNDS, INPUT_SETS),
            "kernel_ms": {k: round(v, 5) for k, v in kern_ms.items()},
            "kernel_time_shares": shares,
            "serial_kernel_sum_ms": tot_kernel_ms,
        }

        # ---- end to end through the public module API with HOST buffers ---------------------------
        # One set of pinned host buffers / device staging buffers per in-flight slot: step i uploads its
        # inputs, runs the hook bodies and downloads features_q + bit maps on slot stream i % nslot;
        # before a slot is reused the host waits for that slot's previous step (the consumer has its
        # result), so uploads of one step overlap downloads of the previous one (PCIe is full duplex).
        host_in = [[torch.empty((B, C, H, Wd), dtype=tdtype).pin_memory() for C, H, Wd in shapes] for _ in range(nslot)]
        for hs in host_in:
            for h, d in zip(hs, sets[0]):
                h.copy_(d)
        host_out = [[torch.empty_like(h).pin_memory() for h in hs] for hs in host_in]
        host_bits = [[None] * len(shapes) for _ in range(nslot)]
        dev_in = [[torch.empty_like(d) for d in sets[0]] for _ in range(nslot)]
        done_ev = [None] * nslot

        def e2e_step(i):
            sl = i % nslot
            if done_ev[sl] is not None:
                done_ev[sl].synchronize()                      # the caller has read this slot's previous result
            with torch.cuda.stream(slot_streams[sl]):
                for h, d in zip(host_in[sl], dev_in[sl]):
                    d.copy_(h, non_blocking=True)
                recs = step(dev_in[sl], sl)
                for k, r in enumerate(recs):
                    host_out[sl][k].copy_(r["features_q"], non_blocking=True)
                    if host_bits[sl][k] is None:
                        host_bits[sl][k] = torch.empty(r["bit_map"].shape, dtype=torch.float32).pin_memory()
                    host_bits[sl][k].copy_(r["bit_map"], non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(slot_streams[sl])
                done_ev[sl] = ev

        def e2e_drain():
            for ev in done_ev:
                if ev is not None:
                    ev.synchronize()

        for i in range(2 * nslot):
            e2e_step(i)
        e2e_drain()
        nsteps_e2e = min(args.steps, 20)
        nsteps_e2e -= nsteps_e2e % nslot
        barrier()
        t0 = time.perf_counter()
        for i in range(nsteps_e2e):
            e2e_step(i)
        e2e_drain()
        wall = time.perf_counter() - t0
        barrier()
        te = torch.tensor([wall], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        e2e_value = world * B * nsteps_e2e / float(te.item())
        h2d = sum(h.numel() * h.element_size() for h in host_in[0])
        d2h = sum(h.numel() * h.element_size() for h in host_out[0]) + sum(hb.numel() * 4 for hb in host_bits[0])

    # ---- multi-GPU parity, outside every timed region: (1) the ranges K2 merged over NVLink peer memory equal an
    #      NCCL all_reduce(MIN) of the ranks' local ranges; (2) this rank's y of a sharded step is bit-identical to
    #      its slice of the UNSHARDED batch recomputed locally from the all-gathered inputs (quantization.py:650-654)
    parity = None
    if world > 1 and not args.unfused and not args.frozen_ranges:
        with torch.no_grad():
            feats = sets[0]
            ok = 1
            why = []
            recs = sharded.run(feats) if sharded is not None else hots[0].run(feats)
            torch.cuda.synchronize()
            if sharded is None and hots[0].xchg[0] is not None:
                for si, (x, (C, H, Wd)) in enumerate(zip(feats, shapes)):
                    s_, a_, keys = ops.reduce_planes(x)
                    local = ops.ranges_decode(keys)
                    want = local.clone()
                    dist.all_reduce(want, op=dist.ReduceOp.MIN)
                    ops._call("mcaq_ranges_reset", keys.data_ptr(), C, ops._stream())
                    ops.reduce_planes_into(x, s_, a_, keys)
                    r_ = ops.morph_fused(s_, a_, C, grid, cm, mp_, KC.pack_soft_mask(quantizers[si].soft_mask), 1.0,
                                         keys=keys, xchg=hots[0].xchg[si])
                    torch.cuda.synchronize()
                    if not torch.equal(r_["packed"], want):
                        ok = 0
                        why.append("C%d merged ranges != all_reduce(MIN)" % (3 + si))
                    hots[0].xchg[si].check()
            full = []
            for x in feats:
                g_ = torch.empty((world * B,) + tuple(x.shape[1:]), device=dev, dtype=x.dtype)
                dist.all_gather_into_tensor(g_, x.contiguous())
                full.append(g_)
            plain = FusedHotPath(analyzer, mapper, quantizers, temperature=1.0, streams=False)
            ref_recs = plain.run(full)
            torch.cuda.synchronize()
            for si, (a_, b_) in enumerate(zip(recs, ref_recs)):
                if not torch.equal(a_["features_q"], b_["features_q"][rank * B:(rank + 1) * B]):
                    ok = 0
                    why.append("C%d y != unsharded recompute" % (3 + si))
                if not torch.equal(a_["bit_map"], b_["bit_map"][rank * B:(rank + 1) * B]):
                    ok = 0
                    why.append("C%d bit map != unsharded recompute" % (3 + si))
            flag = torch.tensor([ok], device=dev)
            dist.all_reduce(flag, op=dist.ReduceOp.MIN)
            parity = "ok" if int(flag.item()) == 1 else ("FAILED on some rank" + (": " + "; ".join(why) if why else ""))
            del full, ref_recs

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32 arithmetic on %s feature maps" % dtype_name, "data": "synthetic",
        "config": {"workload": args.workload, "per_gpu_batch": B, "shapes_CHW": shapes, "grid": grid, "bits": "2-8",
                   "mapper": "MLP (fixture weights)", "soft_mask": True,
                   "ranges": "frozen after calibration (no exchange between ranks)" if args.frozen_ranges else "dynamic per batch"
                   + ("" if world == 1 else (" (all-reduced MIN over ranks, NCCL)" if sharded is not None else
                                             " (min over ranks inside K2 through NVLink peer memory, no collective launch)")),
                   "l2": "%d rotating input sets (%.0f MB) > 126 MB L2, no flush" % (INPUT_SETS, INPUT_SETS * esize * elems * B / 1e6),
                   "launch": ("cuda-graph replay" if graphs is not None else "eager")
                   + (", module-by-module" if args.unfused else ", fused K1/K2/K3 per scale")
                   + ("" if args.no_streams or args.unfused or sharded is not None else ", one stream per scale")
                   + (", %d steps in flight (alternating streams, one workspace each)" % nslot if nslot > 1 else "")
                   + (", ranges all-reduce (NCCL, eager, side stream) between graph segments" if sharded is not None else "")},
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "steps": nsteps_e2e, "api": "FusedHotPath.run(feats) (the hook bodies install() registers), pinned host "
                "buffers, H2D + hooks + D2H per step, %d slot(s) double-buffered, host wall clock" % nslot,
                "per_gpu": e2e_value / world, "host": host_info},
        "gpu_launches": launches_per_step * args.steps,
        "gpu_launches_per_step": launches_per_step,
        "roofline": roofline,
    }
    if parity is not None:
        line["parity_check"] = parity
        line["parity_check_what"] = ("merged ranges == NCCL all_reduce(MIN) of the local ranges; y and bit maps of a sharded "
                                     "step == this rank's slice of the unsharded batch (all-gathered inputs), bit for bit")
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        if _reference_available():
            t_one = _reference_worker((shapes, 2, 0, grid))             # imports + first call
            n, t0 = 0, time.perf_counter()
            while time.perf_counter() - t0 < args.cpu_seconds:
                _reference_worker((shapes, 2, 1 + n, grid))
                n += 1
            dt = time.perf_counter() - t0
            line["cpu_baseline"] = {"value": 2 * n / dt, "unit": UNIT, "cores": 1, "kind": "reference",
                                    "sample": f"{2 * n} images (2-image fp32 batches, same shapes), {dt:.1f} s, the reference's "
                                              "own torch modules on one host core"}
        else:
            v, nimg, dt = cpu_sample(shapes, grid, args.cpu_seconds, 1)
            line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": 1, "kind": "port",
                                    "sample": f"{nimg} images (2-image fp32 batches, same shapes), {dt:.1f} s, numpy oracle"}
    if rank == 0 and world == 1 and args.workload == "yolov8n_640_b64_bf16" and not args.no_secondary:
        # the larger maps of BASELINE configs[2] (32 images per GPU, fp32), same code path, measured right after the
        # headline workload in a child process (its own buffers and graphs); only the key figures are kept
        try:
            out = subprocess.run([sys.executable, os.path.abspath(__file__), "--workload", "yolov8s_1280_b32_f32", "--steps", "40",
                                  "--warmup", "5", "--no-cpu-baseline", "--no-secondary"], capture_output=True, text=True, timeout=300)
            sj = json.loads(out.stdout.strip().splitlines()[-1])
            line["secondary"] = {"workload": "yolov8s_1280_b32_f32", "what": "per-GPU share of BASELINE configs[2] (YOLOv8s @ 1280x1280, "
                                 "32 images, fp32 feature maps 128x160x160 / 256x80x80 / 512x40x40), same fused path",
                                 "ms_per_step": sj["ms_per_step"], "value": sj["value"], "unit": UNIT,
                                 "roofline_frac_whole_step": sj["roofline"]["whole_step"]["frac"],
                                 "achieved_GBs": sj["roofline"]["whole_step"]["achieved"],
                                 "serial_hook_ms": sj["roofline"]["serial_hook"]["ms_per_forward"],
                                 "kernel_ms": sj["roofline"]["kernel_ms"], "clocks": sj.get("clocks")}
        except Exception as e:       # the headline line must not depend on it
            line["secondary"] = {"workload": "yolov8s_1280_b32_f32", "error": f"{type(e).__name__}: {e}"[:200]}
    if rank == 0:
        sys.stdout.flush()
        os.write(out_fd, (json.dumps(line) + "\n").encode())
    if world > 1:
        # orderly shutdown; a watchdog guarantees the process exits even if NCCL teardown stalls
        import threading
        threading.Thread(target=lambda: (time.sleep(30), os._exit(0)), daemon=True).start()
        graphs = keep = None
        torch.cuda.synchronize()
        dist.barrier()
        dist.destroy_process_group()


def run_model(args):
    """Whole-model timing on one GPU: the reference's MCAQYOLO wrapper (unmodified, models/mcaq_yolo.py:222-589) on the
    harness' minimal random-init YOLOv8, (a) hooks inactive, (b) the reference's own torch hooks on CUDA, (c) the
    native hooks after modules.install().  value = images/s of (c); the hooks' share is (c) - (a)."""
    import copy
    import torch
    from harness import ref_model
    assert torch.cuda.is_available()
    pkg = ref_model.load()
    if pkg is None:
        print(json.dumps({"metric": METRIC, "unavailable": "no copy of the reference reachable (baseline/_ref or /root/reference)"}))
        return
    from mcaq_yolo.models.mcaq_yolo import MCAQYOLO
    from mcaq_yolo_b200 import modules as M
    from golden_util import weights
    scale, size, B, dtype_name = MODEL_WORKLOADS[args.workload]
    tdtype = torch.bfloat16 if dtype_name == "bf16" else torch.float32
    torch.backends.cudnn.benchmark = True
    torch.manual_seed(0)
    ref = MCAQYOLO(f"yolov8{scale}", pretrained=False, device="cuda").eval()
    W = weights()
    sd = lambda d: {k: torch.as_tensor(v) for k, v in d.items()}      # noqa: E731
    ref.complexity_analyzer.load_state_dict(sd(W["analyzer"]))
    ref.bit_mapper.load_state_dict(sd(W["mapper"]))
    for q in ref.quantizers.values():
        q.load_state_dict(sd(W["quantizer"]))
    nat = copy.deepcopy(ref)
    layers = list(nat.model.model)
    for idx in nat.backbone_out_indices:
        layers[idx]._forward_hooks.clear()
    nat._mcaq_hooks = [layers[i].register_forward_hook(nat._make_mcaq_hook(i)) for i in nat.backbone_out_indices]
    M.install(nat)
    nat.model.to(tdtype)
    # the reference cannot run a bf16-weight model (its eval quantiser returns fp32 = bf16 * fp32 mask, which the next
    # bf16 layer rejects): its arm keeps fp32 weights and runs under autocast, as its own trainer does (train.py:748)
    ref_ctx = (lambda: torch.autocast("cuda", dtype=tdtype)) if tdtype != torch.float32 else (lambda: torch.autocast("cuda", enabled=False))
    xs = [torch.rand(B, 3, size, size, device="cuda").to(tdtype) for _ in range(3)]

    def timed(fn, n, warm=3):
        with torch.no_grad():
            for i in range(warm):
                fn(xs[i % 3])
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for i in range(n):
                fn(xs[i % 3])
            b.record()
            torch.cuda.synchronize()
        return a.elapsed_time(b) / n

    n = max(3, args.steps)
    t_plain = timed(lambda x: nat.model(x), n)                       # hooks inactive: the bare network
    t_nat = timed(lambda x: nat(x), n)
    def ref_fwd(x):
        with ref_ctx():
            return ref(x.float() if tdtype != torch.float32 else x)
    t_ref = timed(ref_fwd, max(2, min(n, 5)), warm=1)
    # agreement of the bit maps (not timed): the SAME raw backbone features (fp32, TF32 off) through the reference's
    # modules on CUDA and through the native modules, per scale
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    feats = {}
    hs = [ref.model.model[i].register_forward_hook(lambda m_, i_, o, k=i: feats.__setitem__(k, o.detach()))
          for i in ref.backbone_out_indices]
    with torch.no_grad():
        ref.model(xs[0].float()[: min(B, 8)])
        for h in hs:
            h.remove()
        agree = []
        for k in ref.backbone_out_indices:
            b_ref = ref.bit_mapper(ref.complexity_analyzer(feats[k]), 1.0)
            b_nat = nat.bit_mapper(nat.complexity_analyzer(feats[k]), 1.0)
            agree.append(float((b_ref == b_nat).float().mean()))
        _, a_nat = nat(xs[0])
    line = {"metric": "MCAQ-YOLO whole forward (backbone + MCAQ hooks + neck + head), images/s", "value": B / t_nat * 1e3,
            "unit": UNIT, "n_gpus": 1, "steps": n, "warmup": 3, "ms_per_step": t_nat, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": dtype_name, "data": "synthetic images, random-init weights",
            "config": {"workload": args.workload, "model": f"minimal YOLOv8{scale} (tests/harness), {size}x{size}, batch {B}",
                       "wrapper": "the reference's MCAQYOLO, unmodified, with an ultralytics stand-in; native modules via install()"},
            "ms_forward_plain_network": t_plain, "ms_forward_native_hooks": t_nat, "ms_hooks_native": t_nat - t_plain,
            "ms_forward_reference_hooks_cuda": t_ref, "ms_hooks_reference_cuda": t_ref - t_plain,
            "hooks_speedup_vs_reference_on_same_gpu": (t_ref - t_plain) / max(t_nat - t_plain, 1e-9),
            "bit_map_agreement_with_reference_cuda": agree,
            "bit_map_agreement_note": "fraction of tiles with the same width when the reference's modules (torch ON CUDA, fp32, TF32 "
                                      "off) and the native modules see the SAME raw C3/C4/C5 features of the first 8 images",
            "avg_bits": float(a_nat["avg_bits"])}
    print(json.dumps(line), flush=True)


if __name__ == "__main__":
    a = parse()
    if a.workload in MODEL_WORKLOADS:
        run_model(a)
    elif a.impl == "reference":
        run_reference(a)
    else:
        run_native(a)
