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


class ScaleWorkspace:
    """Per-scale persistent state of the fused path: the range keys K1 accumulates into.  They are
    armed once here; afterwards K2's first CTA decodes and re-arms them every step."""

    def __init__(self, C: int, device):
        self.C = C
        self.keys = torch.empty((2 * C,), device=device, dtype=torch.int32)
        ops._call("mcaq_ranges_reset", self.keys.data_ptr(), C, ops._stream())


def fused_scale_forward(feat: torch.Tensor, analyzer, mapper, quantizer, temperature, ws: ScaleWorkspace | None,
                        layer: int = -1) -> dict:
    """Eval-mode hook body for one scale in three launches.  Returns the aux record."""
    B, C, H, W = feat.shape
    x = feat if feat.is_contiguous() else feat.contiguous()
    frozen = quantizer._is_frozen() and quantizer.running_min is not None
    sync = (quantizer.sync_ranges and torch.distributed.is_available() and torch.distributed.is_initialized()
            and torch.distributed.get_world_size(quantizer.process_group) > 1)
    need_ranges = not frozen
    if ws is None or ws.C != C:
        ws = ScaleWorkspace(C, x.device) if need_ranges else None
    s = torch.empty((B, H, W), device=x.device, dtype=torch.float32)
    a = torch.empty((B, H, W), device=x.device, dtype=torch.float32)
    ops._call("mcaq_reduce_planes", x.data_ptr(), ops._dtype_code(x), B, C, H, W, s.data_ptr(), a.data_ptr(),
              ws.keys.data_ptr() if need_ranges else None, ops._stream())
    linear = isinstance(mapper, M.LinearBitMapper)
    sm = quantizer.soft_mask if quantizer.smooth_transitions else None
    r = ops.morph_fused(s, a if sm is not None else None, C, analyzer.grid_size,
                        K.pack_complexity_mlp(analyzer.complexity_mlp),
                        None if linear else K.pack_mapping_network(mapper.mapping_network),
                        None if sm is None else K.pack_soft_mask(sm),
                        temperature, False, ws.keys if need_ranges else None,
                        mapper.min_bits, mapper.max_bits, getattr(mapper, "eps_spread", 1e-3))
    if frozen:
        y = ops.tile_quantize_ranges(x, r["bit_map"], None, quantizer.running_min, quantizer.running_max, r["mask"])
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

    def __init__(self, analyzer, mapper, quantizers, temperature: float = 1.0, streams: bool = True):
        self.analyzer, self.mapper, self.quantizers = analyzer, mapper, list(quantizers)
        self.temperature = temperature
        self.ws = [None] * len(self.quantizers)
        self.use_streams = streams
        self.side = None

    @torch.no_grad()
    def run(self, feats):
        n = len(feats)
        if not self.use_streams or n == 1:
            out = []
            for i, x in enumerate(feats):
                rec, self.ws[i] = fused_scale_forward(x, self.analyzer, self.mapper, self.quantizers[i],
                                                      self.temperature, self.ws[i], i)
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
                                                         self.temperature, self.ws[i], i)
                ev = torch.cuda.Event()
                ev.record(st)
                joins.append(ev)
        out[0], self.ws[0] = fused_scale_forward(feats[0], self.analyzer, self.mapper, self.quantizers[0],
                                                 self.temperature, self.ws[0], 0)
        for ev in joins:
            cur.wait_event(ev)
        if not torch.cuda.is_current_stream_capturing():
            for rec in out[1:]:          # produced on a side stream, consumed on the caller's stream
                for k in ("complexity", "bit_map", "features_q"):
                    rec[k].record_stream(cur)
        return out
