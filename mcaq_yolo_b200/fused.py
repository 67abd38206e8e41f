"""Three-launch inference hook: K1 channel sweep -> K2 fused morphology/nets -> K3 quantize.

`FusedMcaqHook` is what `install(model, fused=True)` registers in place of the reference's hook
closure (models/mcaq_yolo.py:402-457): it reads the same `model._mcaq_state` protocol, produces
the same aux records and uses the same three module objects for their parameters/buffers, but
runs the eval path as three kernels per scale with no host synchronisation.  Training,
calibration, user-supplied mappers and `normalize_complexity` fall back to the module-by-module
path (`modules.mcaq_hook_forward`), which has identical semantics.
"""
from __future__ import annotations

import torch

from . import constants as K
from . import modules as M
from . import ops


def run_forked(tasks, side_streams):
    """Run independent callables concurrently: task 0 on the current stream, task i on
    side_streams[i-1], fork/join with events (works eagerly and under CUDA-graph capture).
    Returns the list of results."""
    n = len(tasks)
    if n == 1 or not side_streams:
        return [t() for t in tasks]
    cur = torch.cuda.current_stream()
    fork = torch.cuda.Event()
    fork.record(cur)
    out = [None] * n
    joins = []
    for i in range(1, n):
        st = side_streams[i - 1]
        st.wait_event(fork)
        with torch.cuda.stream(st):
            out[i] = tasks[i]()
            ev = torch.cuda.Event()
            ev.record(st)
            joins.append(ev)
    out[0] = tasks[0]()
    for ev in joins:
        cur.wait_event(ev)
    return out


def _record_on(cur, obj):
    """Tensors produced on a side stream and consumed on `cur`: tell the allocator (eager mode)."""
    if torch.cuda.is_current_stream_capturing():
        return
    if torch.is_tensor(obj):
        obj.record_stream(cur)
    elif isinstance(obj, dict):
        for v in obj.values():
            _record_on(cur, v)
    elif isinstance(obj, (list, tuple)):
        for v in obj:
            _record_on(cur, v)


class ScaleWorkspace:
    """Per-scale persistent state of the fused path: the range keys K1 accumulates into.  They are
    armed once here; afterwards K2's first CTA decodes and re-arms them every step."""

    def __init__(self, C: int, device):
        self.C = C
        self.keys = torch.empty((2 * C,), device=device, dtype=torch.int32)
        ops._call("mcaq_ranges_reset", self.keys.data_ptr(), C, ops._stream())


import os as _os
_K1_PRIO = bool(int(_os.environ.get("MCAQ_K1_PRIO", "0")))      # tuning aid: K1 on the high-priority stream too


def mapper_block(mapper, temperature) -> torch.Tensor:
    """Parameter block of the MLP mapper for K2: with its step table when the network is monotone in c
    (constants.mapping_is_monotone), the plain block -- K2 then evaluates the network per tile -- when it
    is not (enforce_monotonicity=False, or negative weights before enforce_weight_constraints())."""
    if K.mapping_is_monotone(mapper):
        return K.pack_mapping_steps(mapper.mapping_network, temperature, mapper.min_bits, mapper.max_bits)
    return K.pack_mapping_network(mapper.mapping_network)


def fused_scale_forward(feat: torch.Tensor, analyzer, mapper, quantizer, temperature, ws: ScaleWorkspace | None,
                        layer: int = -1, xchg=None, k2_stream=None) -> dict:
    """Eval-mode hook body for one scale in three launches.  Returns the aux record.
    xchg: a peer.RangeExchange when the batch is sharded over the GPUs of a node -- the range merge
    then happens inside K2 / K3 over peer memory instead of a collective between them."""
    B, C, H, W = feat.shape
    x = ops.as_kernel_layout(feat)            # NCHW-contiguous or channels_last (native NHWC K1 / K3)
    frozen = quantizer._is_frozen() and quantizer.running_min is not None
    sync = (quantizer.sync_ranges and torch.distributed.is_available() and torch.distributed.is_initialized()
            and torch.distributed.get_world_size(quantizer.process_group) > 1)
    need_ranges = not frozen
    if ws is None or ws.C != C:
        ws = ScaleWorkspace(C, x.device) if need_ranges else None
    s = torch.empty((B, H, W), device=x.device, dtype=torch.float32)
    a = torch.empty((B, H, W), device=x.device, dtype=torch.float32)
    if k2_stream is not None and _K1_PRIO:
        cur0 = torch.cuda.current_stream()
        ev0 = torch.cuda.Event()
        ev0.record(cur0)
        k2_stream.wait_event(ev0)
        with torch.cuda.stream(k2_stream):
            ops.reduce_planes_into(x, s, a, ws.keys if need_ranges else None)
            ev1 = torch.cuda.Event()
            ev1.record(k2_stream)
        cur0.wait_event(ev1)
    else:
        ops.reduce_planes_into(x, s, a, ws.keys if need_ranges else None)
    linear = isinstance(mapper, M.LinearBitMapper)
    sm = quantizer.soft_mask if quantizer.smooth_transitions else None
    def k2():
        return ops.morph_fused(s, a if sm is not None else None, C, analyzer.grid_size,
                               K.pack_complexity_mlp(analyzer.complexity_mlp),
                               None if linear else mapper_block(mapper, temperature),
                               None if sm is None else K.pack_soft_mask(sm),
                               temperature, False, ws.keys if need_ranges else None,
                               mapper.min_bits, mapper.max_bits, getattr(mapper, "eps_spread", 1e-3),
                               xchg=xchg if need_ranges else None)
    if k2_stream is None:
        r = k2()
    else:
        # the latency-bound morphology kernel on a HIGH-PRIORITY stream (fork / join by events, capturable): its few
        # long-lived CTAs are placed ahead of the pending CTAs of the bandwidth kernels of other scales / steps
        cur = torch.cuda.current_stream()
        ev = torch.cuda.Event()
        ev.record(cur)
        k2_stream.wait_event(ev)
        with torch.cuda.stream(k2_stream):
            r = k2()
            done = torch.cuda.Event()
            done.record(k2_stream)
        cur.wait_event(done)
        _record_on(cur, r)
        if not torch.cuda.is_current_stream_capturing():
            s.record_stream(k2_stream)
            a.record_stream(k2_stream)
    if frozen:
        y = ops.tile_quantize_ranges(x, r["bit_map"], None, quantizer.running_min, quantizer.running_max, r["mask"])
    elif xchg is not None and xchg.world > 1:
        # K2's first CTA published this rank's ranges at its start and merged all ranks' at its end
        y = ops.tile_quantize_ranges(x, r["bit_map"], r["packed"], None, None, r["mask"])
    else:
        packed = r["packed"]
        if sync:
            M.allreduce_ranges(packed, quantizer.process_group)
        y = ops.tile_quantize_ranges(x, r["bit_map"], packed, None, None, r["mask"])
    return {"layer": layer, "complexity": r["complexity"], "bit_map": r["bit_map"], "features_q": y}, ws


def _fusable(analyzer, mapper, quantizer, training: bool, calibrating: bool, normalize: bool) -> bool:
    if training or calibrating or normalize or torch.is_grad_enabled():
        return False
    if not isinstance(analyzer, M.MorphologicalComplexityAnalyzer) or not isinstance(quantizer, M.SpatialAdaptiveQuantization):
        return False
    if isinstance(mapper, M.LinearBitMapper):
        return True
    return isinstance(mapper, M.ComplexityToBitMappingNetwork) and mapper._kernel_ok()


class FusedMcaqHook:
    """Forward hook with the reference's protocol, one instance per hooked layer."""

    def __init__(self, model, layer_idx: int):
        self.model = model
        self.layer_idx = layer_idx
        self.ws = None
        from . import _lib
        _lib.load().mcaq_morph_policy(1)      # hooks run serially inside the backbone: latency policy

    def __call__(self, module, inputs, output):
        model = self.model
        state = model._mcaq_state
        if not state.get("active", False):
            return None
        if not torch.is_tensor(output) or output.dim() != 4:
            return None
        quantize = state.get("quantize", True)
        calibrating = state.get("calibrating", False)
        analyzer, mapper = model.complexity_analyzer, model.bit_mapper
        quantizer = model.quantizers[str(self.layer_idx)]
        normalize = getattr(model, "normalize_complexity", False)
        if quantize and _fusable(analyzer, mapper, quantizer, model.training, calibrating, normalize):
            rec, self.ws = fused_scale_forward(output, analyzer, mapper, quantizer, state.get("temperature", 1.0),
                                               self.ws, self.layer_idx)
        else:
            rec = M.mcaq_hook_forward(output, analyzer, mapper, quantizer, state.get("temperature", 1.0),
                                      quantize=quantize, training=model.training, calibrating=calibrating,
                                      normalize_complexity=normalize, layer=self.layer_idx)
        state["aux"].append(rec)
        return rec["features_q"] if quantize else None


class FusedHotPath:
    """The three hooks of one forward as a stand-alone object (bench.py, serving without the model
    wrapper): `run(feats)` takes the C3/C4/C5 maps and returns the aux records."""

    def __init__(self, analyzer, mapper, quantizers, temperature: float = 1.0, streams: bool = True,
                 exchanges=None, latency: bool = False, k2_priority: bool = False):
        self.analyzer, self.mapper, self.quantizers = analyzer, mapper, list(quantizers)
        self.temperature = temperature
        self.ws = [None] * len(self.quantizers)
        self.use_streams = streams
        self.side = None
        # optional: one high-priority stream per scale for the morphology kernel (bench.py --k2-priority)
        self.k2s = [torch.cuda.Stream(priority=-1) for _ in self.quantizers] if k2_priority else [None] * len(self.quantizers)
        # one peer.RangeExchange per scale when the batch is sharded over the GPUs of a node
        self.xchg = list(exchanges) if exchanges is not None else [None] * len(self.quantizers)
        # latency=True: a serial caller -- split every image over a cluster (library-wide policy switch)
        from . import _lib
        _lib.load().mcaq_morph_policy(1 if latency else 0)

    @torch.no_grad()
    def run(self, feats):
        n = len(feats)
        if not self.use_streams or n == 1:
            out = []
            for i, x in enumerate(feats):
                rec, self.ws[i] = fused_scale_forward(x, self.analyzer, self.mapper, self.quantizers[i],
                                                      self.temperature, self.ws[i], i, self.xchg[i], self.k2s[i])
                out.append(rec)
            return out
        # scales are independent: fork one stream per extra scale so K2's per-image latency of one
        # scale overlaps the bandwidth kernels of the others (graph-capturable fork/join)
        cur = torch.cuda.current_stream()
        if self.side is None or len(self.side) != n - 1:
            self.side = [torch.cuda.Stream() for _ in range(n - 1)]
        fork = torch.cuda.Event()
        fork.record(cur)
        out = [None] * n
        joins = []
        for i in range(1, n):
            st = self.side[i - 1]
            st.wait_event(fork)
            with torch.cuda.stream(st):
                out[i], self.ws[i] = fused_scale_forward(feats[i], self.analyzer, self.mapper, self.quantizers[i],
                                                         self.temperature, self.ws[i], i, self.xchg[i], self.k2s[i])
                ev = torch.cuda.Event()
                ev.record(st)
                joins.append(ev)
        out[0], self.ws[0] = fused_scale_forward(feats[0], self.analyzer, self.mapper, self.quantizers[0],
                                                 self.temperature, self.ws[0], 0, self.xchg[0], self.k2s[0])
        for ev in joins:
            cur.wait_event(ev)
        if not torch.cuda.is_current_stream_capturing():
            for rec in out[1:]:          # produced on a side stream, consumed on the caller's stream
                for k in ("complexity", "bit_map", "features_q"):
                    rec[k].record_stream(cur)
        return out


class ShardedHotPath:
    """Batch-sharded variant for world_size > 1 (one process per GPU, images split by rank).

    The only exchange the inference path needs is the per-channel range merge.  It is
    kept OUT of the captured graphs and issued as one eager NCCL all-reduce(MIN) over a single
    buffer holding all scales' packed ranges, on a side stream, overlapped with K2:

        graph A : K1 + decode + re-arm keys, all scales          (ranges of this rank's shard)
        eager   : side stream waits for A, all_reduce(MIN) of sum_s 2*C_s floats
        graph B1: K2 for all scales (does not need the ranges)    <- overlaps the collective
        eager   : main stream waits for the all-reduce
        graph B2: K3 for all scales with the merged ranges
    """

    def __init__(self, analyzer, mapper, quantizers, shapes, device, temperature: float = 1.0, group=None):
        self.analyzer, self.mapper, self.quantizers = analyzer, mapper, list(quantizers)
        self.temperature, self.group = temperature, group
        self.C = [c for c, _, _ in shapes]
        self.off = [0]
        for c in self.C:
            self.off.append(self.off[-1] + 2 * c)
        self.packed_all = torch.empty((self.off[-1],), device=device, dtype=torch.float32)
        self.ws = [ScaleWorkspace(c, device) for c in self.C]
        self.comm_stream = torch.cuda.Stream(device=device)
        self.side = [torch.cuda.Stream(device=device) for _ in range(len(self.C) - 1)]

    def packed(self, i):
        return self.packed_all[self.off[i]:self.off[i + 1]]

    @torch.no_grad()
    def sweep(self, feats):
        """Phase A: K1 per scale (one stream each), ranges decoded into the shared buffer, keys re-armed."""
        def one(i, x):
            B, C, H, W = x.shape
            s = torch.empty((B, H, W), device=x.device, dtype=torch.float32)
            a = torch.empty((B, H, W), device=x.device, dtype=torch.float32)
            keys = self.ws[i].keys
            ops._call("mcaq_reduce_planes", x.data_ptr(), ops._dtype_code(x), B, C, H, W, s.data_ptr(),
                      a.data_ptr(), keys.data_ptr(), ops._stream())
            ops._call("mcaq_ranges_decode", keys.data_ptr(), C, self.packed(i).data_ptr(), ops._stream())
            ops._call("mcaq_ranges_reset", keys.data_ptr(), C, ops._stream())
            return (s, a)
        planes = run_forked([lambda i=i, x=x: one(i, x) for i, x in enumerate(feats)], self.side)
        _record_on(torch.cuda.current_stream(), planes[1:])
        return planes

    def exchange(self):
        """Eager collective on the side stream; returns the event the quantize phase must wait for."""
        import torch.distributed as dist
        cur = torch.cuda.current_stream()
        ready = torch.cuda.Event()
        ready.record(cur)
        self.comm_stream.wait_event(ready)
        with torch.cuda.stream(self.comm_stream):
            if dist.is_initialized() and dist.get_world_size(self.group) > 1:
                dist.all_reduce(self.packed_all, op=dist.ReduceOp.MIN, group=self.group)
            done = torch.cuda.Event()
            done.record(self.comm_stream)
        return done

    @torch.no_grad()
    def nets(self, feats, planes):
        """Phase B1: K2 per scale (bit maps, masks); independent of the ranges."""
        linear = isinstance(self.mapper, M.LinearBitMapper)

        def one(i, x, s, a):
            q = self.quantizers[i]
            sm = q.soft_mask if q.smooth_transitions else None
            return ops.morph_fused(s, a if sm is not None else None, x.shape[1], self.analyzer.grid_size,
                                   K.pack_complexity_mlp(self.analyzer.complexity_mlp),
                                   None if linear else mapper_block(self.mapper, self.temperature),
                                   None if sm is None else K.pack_soft_mask(sm), self.temperature, False, None,
                                   self.mapper.min_bits, self.mapper.max_bits,
                                   getattr(self.mapper, "eps_spread", 1e-3))
        out = run_forked([lambda i=i, x=x, p=p: one(i, x, p[0], p[1])
                          for i, (x, p) in enumerate(zip(feats, planes))], self.side)
        _record_on(torch.cuda.current_stream(), out[1:])
        return out

    @torch.no_grad()
    def quantize(self, feats, nets_out):
        """Phase B2: K3 per scale with the merged ranges."""
        def one(i, x, r):
            y = ops.tile_quantize_ranges(x, r["bit_map"], self.packed(i), None, None, r["mask"])
            return {"layer": i, "complexity": r["complexity"], "bit_map": r["bit_map"], "features_q": y}
        recs = run_forked([lambda i=i, x=x, r=r: one(i, x, r)
                           for i, (x, r) in enumerate(zip(feats, nets_out))], self.side)
        _record_on(torch.cuda.current_stream(), recs[1:])
        return recs

    def run(self, feats):
        """Eager composition of the four phases (graphs: see bench.py)."""
        planes = self.sweep(feats)
        done = self.exchange()
        nets_out = self.nets(feats, planes)
        torch.cuda.current_stream().wait_event(done)
        return self.quantize(feats, nets_out)
