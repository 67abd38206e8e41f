"""Host-side mirror of the reference's interface for the hot path.

Same class names, constructor arguments, call signatures, error behaviour and state_dict keys as
the reference's `mcaq_yolo/core/{morphology,bit_allocation,quantization}.py`, so the reference's
forward hook (`models/mcaq_yolo.py:409-455`) and checkpoints work unchanged; the arithmetic runs
in the sm_100a kernels of libmcaq_b200.so.  Inference (no grad) is all kernels; in training the
HBM-heavy parts (channel sweep, phi, fractional-bit quantise forward/backward) are kernels and
only the three tiny networks (2.9k / 4.6k / 170 parameters on (B,ht,wt) tensors) stay in torch
autograd.  CUDA tensors only: there is no CPU fallback.
"""
from __future__ import annotations

import math
from typing import Optional

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import constants as K
from . import ops
from . import train_nets as TN


# ---------------------------------------------------------------------------------------------
# one channel sweep per feature map, shared by analyzer and quantizer
# ---------------------------------------------------------------------------------------------
class _SweepCache:
    """Result of K1 for the most recent feature map.  Keyed on tensor identity + version (the
    cache keeps the tensor alive, so its address cannot be recycled while the entry exists)."""

    def __init__(self):
        self.x = None
        self.version = -1
        self.value = None

    def get(self, x: torch.Tensor, want_ranges: bool):
        hit = self.x is x and self.version == x._version and self.value is not None
        if hit and (not want_ranges or self.value[2] is not None):
            return self.value
        self.value = ops.reduce_planes(x, want_ranges=want_ranges)
        self.x, self.version = x, x._version
        return self.value

    def clear(self):
        self.x, self.value, self.version = None, None, -1


_SWEEP = _SweepCache()


def allreduce_ranges(packed: torch.Tensor, group=None) -> torch.Tensor:
    """Merge per-rank channel ranges: `packed` = [min_c..., -max_c...], so ONE MIN all-reduce of
    2C floats yields the ranges of the whole (sharded) batch -- the only collective the inference
    path needs, and none at all once calibration is frozen."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(packed, op=dist.ReduceOp.MIN, group=group)
    return packed


def shard_batch(n: int, rank: int, world: int):
    """Contiguous batch slice of rank `rank` (images are independent)."""
    base, rem = divmod(n, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def channel_sweep(x: torch.Tensor, want_ranges: bool = True):
    """(sum_c x, sum_c |x|, range keys) of a (B,C,H,W) map: one HBM pass, reused by the next
    consumer of the same tensor."""
    return _SWEEP.get(x, want_ranges)


# ---------------------------------------------------------------------------------------------
# analyzer                                                      reference: core/morphology.py:17
# ---------------------------------------------------------------------------------------------
class MorphologicalComplexityAnalyzer(nn.Module):
    def __init__(self, grid_size: int = 8, device: str = "cuda", metric_backend: str = "gpu",
                 canny_impl: str = "cv2compat", binarize_impl: str = "adaptive",
                 contour_components: bool = True):
        super().__init__()
        if metric_backend != "gpu" or canny_impl != "cv2compat" or binarize_impl != "adaptive" \
                or not contour_components:
            raise NotImplementedError(
                "mcaq_b200 implements the reference's default hot path only (metric_backend='gpu', "
                "canny_impl='cv2compat', binarize_impl='adaptive', contour_components=True); the cv2 / "
                "legacy variants are offline reproduction paths of the reference")
        self.grid_size = grid_size
        self.device = device
        self.metric_backend = metric_backend
        self.canny_impl = canny_impl
        self.binarize_impl = binarize_impl
        self.contour_components = contour_components
        self.complexity_mlp = nn.Sequential(
            nn.Linear(8, 64), nn.LayerNorm(64), nn.ReLU(inplace=True),
            nn.Linear(64, 32), nn.LayerNorm(32), nn.ReLU(inplace=True),
            nn.Linear(32, 1), nn.Sigmoid(),
        ).to(device)
        nn.init.xavier_uniform_(self.complexity_mlp[-2].weight, gain=3.0)
        nn.init.zeros_(self.complexity_mlp[-2].bias)
        self.register_buffer("feature_weights", torch.ones(5, device=device) / 5)

    def _tile_size(self, H: int) -> int:
        raw = max(4, H // self.grid_size)
        return 1 << (raw.bit_length() - 1)

    @torch.no_grad()
    def compute_phi_tiles(self, features: torch.Tensor):
        """(B,C,H,W) -> phi (B,ht,wt,8) and the five named metrics (morphology.py:798-873)."""
        s, _, _ = channel_sweep(features, want_ranges=True)
        phi = ops.morph_phi(s, features.shape[1], self.grid_size, K.device_constants(features.device))
        names = ("fractal", "texture", "gradient", "edge", "contour")
        return phi, {n: phi[..., i] for i, n in enumerate(names)}

    def bilateral_filter(self, complexity_map: torch.Tensor, sigma_spatial: float = 2.0,
                         sigma_range: float = 0.1, kernel_size: int = 5) -> torch.Tensor:
        """Differentiable torch bilateral filter for the training path (tiny (B,ht,wt) maps;
        morphology.py:309-354): all k*k neighbours at once, so a handful of launches per call."""
        B, H, W = complexity_map.shape
        r = kernel_size // 2
        nb = F.unfold(F.pad(complexity_map.unsqueeze(1), (r, r, r, r), mode="replicate"), kernel_size)
        centre = complexity_map.reshape(B, 1, H * W)
        ax = torch.arange(kernel_size, device=complexity_map.device, dtype=torch.float32) - r
        dist2 = ax.view(-1, 1) ** 2 + ax.view(1, -1) ** 2
        w_space = torch.exp(-dist2 / (2 * sigma_spatial ** 2)).reshape(1, -1, 1)
        wgt = w_space * torch.exp(-((nb - centre) ** 2) / (2 * sigma_range ** 2))
        return ((wgt * nb).sum(dim=1) / (wgt.sum(dim=1) + 1e-8)).reshape(B, H, W)

    def score_image(self, features: torch.Tensor) -> torch.Tensor:
        """Deterministic Eq.(8) score per image (morphology.py:923-937)."""
        phi, _ = self.compute_phi_tiles(features)
        with torch.no_grad():
            alpha = self.feature_weights.detach().abs()
            alpha = alpha / alpha.sum().clamp(min=1e-8)
            c = (phi[..., :5] * alpha.view(1, 1, 1, 5)).sum(dim=-1)
            return c.mean(dim=(1, 2)).clamp(0.0, 1.0)

    def forward(self, features: torch.Tensor, return_detailed: bool = False):
        phi, detailed = self.compute_phi_tiles(features)
        B, ht, wt, _ = phi.shape
        needs_grad = torch.is_grad_enabled() and any(p.requires_grad for p in self.complexity_mlp.parameters())
        if needs_grad:
            # training: the inference kernel forward (raw MLP output kept), native backward through clamp, bilateral
            # filter and the MLP with its LayerNorms (csrc/train_nets.cu) -- two launches instead of ~150
            flat = TN.flat_params(self.complexity_mlp.parameters(), owner=self.complexity_mlp)
            cmap = TN.ComplexityTrainFn.apply(phi, flat, K.pack_complexity_mlp(self.complexity_mlp),
                                              K.device_constants(features.device))
        else:
            cmap = ops.complexity(phi, K.pack_complexity_mlp(self.complexity_mlp),
                                  K.device_constants(features.device))
        return (cmap, detailed) if return_detailed else cmap


# ---------------------------------------------------------------------------------------------
# bit mappers                                              reference: core/bit_allocation.py:12, 83
# ---------------------------------------------------------------------------------------------
def _normalize_complexity_shape(c: torch.Tensor) -> torch.Tensor:
    if not isinstance(c, torch.Tensor):
        raise TypeError(f"complexity must be torch.Tensor, got {type(c)}")
    if c.dim() == 2:
        return c.unsqueeze(0)
    if c.dim() == 3:
        return c
    if c.dim() == 4:
        return c.mean(dim=1)
    raise ValueError(f"Unsupported complexity dim={c.dim()}, expected 2, 3, or 4.")


def _ste_finish(bits, temperature, lo, hi, continuous):
    if temperature is not None:
        bits = bits * max(float(temperature), 0.1)
    bits = bits + (bits.clamp(lo, hi) - bits).detach()
    if not continuous:
        bits = bits + (torch.round(bits) - bits).detach()
    return bits


class LinearBitMapper(nn.Module):
    def __init__(self, min_bits: int = 2, max_bits: int = 8, eps_spread: float = 1e-3):
        super().__init__()
        self.min_bits = float(min_bits)
        self.max_bits = float(max_bits)
        self.eps_spread = float(eps_spread)

    def enforce_weight_constraints(self):
        """No-op (parameter-free); interface parity with the MLP mapper."""

    def forward(self, complexity: torch.Tensor, temperature: Optional[float] = None,
                return_continuous: bool = False) -> torch.Tensor:
        c = _normalize_complexity_shape(complexity)
        if torch.is_grad_enabled() and c.requires_grad:
            flat = c.reshape(c.shape[0], -1).float()
            lo = torch.quantile(flat, 0.02, dim=1, keepdim=True).unsqueeze(-1)
            hi = torch.quantile(flat, 0.98, dim=1, keepdim=True).unsqueeze(-1)
            spread = hi - lo
            rel = ((c - lo) / (spread + 1e-8)).clamp(0.0, 1.0)
            cn = torch.where(spread > self.eps_spread, rel, c.clamp(0.0, 1.0))
            bits = self.min_bits + (self.max_bits - self.min_bits) * cn
            return _ste_finish(bits, temperature, self.min_bits, self.max_bits, return_continuous)
        return ops.bit_mapper(c, None, temperature, return_continuous, self.min_bits, self.max_bits,
                              self.eps_spread)


class ComplexityToBitMappingNetwork(nn.Module):
    def __init__(self, min_bits: int = 2, max_bits: int = 8, hidden_dims: list = [32, 64, 32],
                 enforce_monotonicity: bool = True):
        super().__init__()
        self.min_bits = float(min_bits)
        self.max_bits = float(max_bits)
        self.enforce_monotonicity = enforce_monotonicity
        self.hidden_dims = list(hidden_dims)
        layers, d = [], 3
        for h in hidden_dims:
            layers += [nn.Linear(d, h), nn.BatchNorm1d(h), nn.ReLU(inplace=True)]
            d = h
        layers += [nn.Linear(d, 1), nn.Sigmoid()]
        self.mapping_network = nn.Sequential(*layers)
        # multi-GPU training: a peer.RangeExchange.create(128) makes the BatchNorm statistics those of the whole
        # (sharded) batch, merged inside the kernels over NVLink peer memory; None = this rank's rows only
        self.stat_exchange = None
        for m in self.mapping_network:
            if isinstance(m, nn.Linear):
                nn.init.xavier_uniform_(m.weight, gain=0.5)
                if enforce_monotonicity:
                    m.weight.data = m.weight.data.abs()
                nn.init.constant_(m.bias, 0.1)

    _normalize_complexity_shape = staticmethod(_normalize_complexity_shape)

    def enforce_weight_constraints(self):
        """W <- |W|, gamma_BN <- |gamma| (Eq.18, bit_allocation.py:186-197)."""
        if self.enforce_monotonicity:
            for m in self.mapping_network.modules():
                if isinstance(m, (nn.Linear, nn.BatchNorm1d)):
                    m.weight.data = m.weight.data.abs()

    def create_augmented_features(self, complexity: torch.Tensor) -> torch.Tensor:
        return torch.cat([complexity, complexity ** 2, torch.log1p(complexity)], dim=-1)

    def _kernel_ok(self) -> bool:
        return (not self.training) and self.hidden_dims == [32, 64, 32]

    def forward(self, complexity: torch.Tensor, temperature: Optional[float] = None,
                return_continuous: bool = False) -> torch.Tensor:
        c = _normalize_complexity_shape(complexity)
        wants_grad = torch.is_grad_enabled() and (
            c.requires_grad or any(p.requires_grad for p in self.mapping_network.parameters()))
        if self._kernel_ok() and not wants_grad:
            return ops.bit_mapper(c, K.pack_mapping_network(self.mapping_network), temperature,
                                  return_continuous, self.min_bits, self.max_bits)
        B, H, W = c.shape
        if self.training and self.hidden_dims == [32, 64, 32] and c.is_cuda:
            # train mode: BatchNorm batch statistics over all B*ht*wt tiles (of all ranks when `stat_exchange`
            # is set), running statistics updated, one cluster launch per direction (csrc/train_nets.cu)
            seq = self.mapping_network
            bits = TN.MapperTrainFn.apply(c.reshape(-1), TN.flat_params(seq.parameters(), owner=seq), (seq[1], seq[4], seq[7]),
                                          temperature, self.min_bits, self.max_bits, self.stat_exchange).reshape(B, H, W)
            if not return_continuous:
                bits = bits + (torch.round(bits) - bits).detach()
            return bits
        # eval-mode BatchNorm with gradients wanted, or a non-default width: torch autograd on the tiny problem
        c = c.clamp(0.0, 1.0)
        h = self.mapping_network(self.create_augmented_features(c.reshape(-1, 1)))
        bits = (self.min_bits + (self.max_bits - self.min_bits) * h).reshape(B, H, W)
        return _ste_finish(bits, temperature, self.min_bits, self.max_bits, return_continuous)


# ---------------------------------------------------------------------------------------------
# soft mask + quantizer                                     reference: core/quantization.py:168, 242
# ---------------------------------------------------------------------------------------------
class LearnedSoftMask(nn.Module):
    def __init__(self, hidden: int = 8, kernel_size: int = 5):
        super().__init__()
        self.net = nn.Sequential(nn.Conv2d(2, hidden, 3, padding=1), nn.ReLU(inplace=True),
                                 nn.Conv2d(hidden, 2, 1))
        nn.init.normal_(self.net[-1].weight, std=1e-3)
        with torch.no_grad():
            self.net[-1].bias.copy_(torch.tensor([4.0, 0.0]))
        k = kernel_size
        ax = torch.arange(k, dtype=torch.float32) - k // 2
        g1 = torch.exp(-ax ** 2 / (2 * (k / 3.0) ** 2))
        g1 = g1 / g1.sum()
        self.register_buffer("smooth_kernel", (g1.unsqueeze(0) * g1.unsqueeze(1)).view(1, 1, k, k))
        self.kernel_size = k
        self.hidden = hidden

    def forward(self, bit_map: torch.Tensor, x: torch.Tensor) -> torch.Tensor:
        """m(p) of shape (B,1,H,W) in [0,1] (quantization.py:213-239)."""
        B, C, H, W = x.shape
        _, abs_plane, _ = channel_sweep(x, want_ranges=False)
        wants_grad = torch.is_grad_enabled() and (
            bit_map.requires_grad or any(p.requires_grad for p in self.net.parameters()))
        if not wants_grad and self.hidden == 8 and self.kernel_size == 5:
            return ops.soft_mask(bit_map, abs_plane, C, K.pack_soft_mask(self)).unsqueeze(1)
        if self.hidden == 8 and self.kernel_size == 5:
            # training: inference kernel forward, native backward (smoothing / upsampling transposed, softmax, the
            # two convolutions) -- gradients to the bit map and the four parameter tensors
            flat = TN.flat_params(self.net.parameters(), owner=self.net)
            return TN.SoftMaskTrainFn.apply(bit_map, flat, abs_plane, C, K.pack_soft_mask(self)).unsqueeze(1)
        Ht, Wt = bit_map.shape[-2:]
        with torch.no_grad():
            act = F.adaptive_avg_pool2d((abs_plane / C).unsqueeze(1), (Ht, Wt))
            act = act / (act.amax(dim=(2, 3), keepdim=True) + 1e-8)
        bits_norm = ((bit_map.unsqueeze(1).float() - 2.0) / 6.0).clamp(0.0, 1.0)
        # full fp32 convolutions (cuDNN's default TF32 would cost ~1e-3 relative on m)
        with torch.backends.cudnn.flags(enabled=True, allow_tf32=False):
            logits = self.net(torch.cat([bits_norm, act], dim=1))
            m = torch.softmax(logits, dim=1)[:, :1]
            m = F.interpolate(m, size=(H, W), mode="nearest")
            p = self.kernel_size // 2
            return F.conv2d(F.pad(m, (p, p, p, p), mode="replicate"), self.smooth_kernel)


class _FractionalQuant(torch.autograd.Function):
    """y = m * [(1-f) Q_floor(b)(x) + f Q_floor(b)+1(x)] with the STE backward of
    quantization.py:69-118, 699-727 as two kernels."""

    @staticmethod
    def forward(ctx, x, bit_map, mask, qtable):
        m3 = None if mask is None else mask.reshape(mask.shape[0], mask.shape[-2], mask.shape[-1])
        y = ops.tile_quantize_train_fwd(x, bit_map, qtable, m3)
        ctx.save_for_backward(x, bit_map, qtable, *(() if m3 is None else (m3,)))
        ctx.has_mask = m3 is not None
        ctx.mask_shape = None if mask is None else mask.shape
        return y

    @staticmethod
    def backward(ctx, gy):
        saved = ctx.saved_tensors
        x, bit_map, qtable = saved[:3]
        m3 = saved[3] if ctx.has_mask else None
        dx, dbit, dmask = ops.tile_quantize_train_bwd(gy, x, bit_map, qtable, m3)
        if dmask is not None:
            dmask = dmask.reshape(ctx.mask_shape)
        return dx, dbit.to(bit_map.dtype), dmask, None


class _FractionalQuantKD(torch.autograd.Function):
    """(y, mse(y, teacher)) in one forward kernel and one backward kernel: the feature-level
    distillation term of train.py:599-610 folded into the fractional-bit quantiser, so the loss
    never re-reads y (forward) and its gradient 2 (y - t) / N is formed in registers (backward)."""

    @staticmethod
    def forward(ctx, x, bit_map, mask, qtable, teacher):
        m3 = None if mask is None else mask.reshape(mask.shape[0], mask.shape[-2], mask.shape[-1])
        y, kd_sum = ops.tile_quantize_train_fwd_kd(x, bit_map, qtable, m3, teacher)
        ctx.save_for_backward(x, bit_map, qtable, teacher, *(() if m3 is None else (m3,)))
        ctx.has_mask = m3 is not None
        ctx.mask_shape = None if mask is None else mask.shape
        return y, (kd_sum / y.numel()).float()

    @staticmethod
    def backward(ctx, gy, gmse):
        saved = ctx.saved_tensors
        x, bit_map, qtable, teacher = saved[:4]
        m3 = saved[4] if ctx.has_mask else None
        coef = gmse.float() * (2.0 / x.numel())                 # device scalar: no host round trip
        dx, dbit, dmask = ops.tile_quantize_train_bwd_kd(gy, x, bit_map, qtable, m3, teacher, coef)
        if dmask is not None:
            dmask = dmask.reshape(ctx.mask_shape)
        return dx, dbit.to(bit_map.dtype), dmask, None, None


class SpatialAdaptiveQuantization(nn.Module):
    """Eq.19  X_q(p) = m(p) * Q_{b_T(p)}(X(p))  (quantization.py:242-754), `minmax` calibration,
    per-channel ranges.  `sync_ranges=True` (opt-in; the reference has no collective, so the default
    keeps a forward hook free of one) all-reduces the per-channel ranges over `process_group` so that a
    batch sharded over ranks quantises exactly like the unsharded reference batch; every rank must then
    run the same number of forwards."""

    def __init__(self, calibration_mode: str = "minmax", smooth_transitions: bool = True,
                 per_channel: bool = True, learned_rounding: bool = False, momentum: float = 0.99,
                 process_group=None, sync_ranges: bool = False):
        super().__init__()
        if calibration_mode != "minmax" or not per_channel or learned_rounding:
            raise NotImplementedError(
                "mcaq_b200 implements what MCAQYOLO instantiates (models/mcaq_yolo.py:466-470): "
                "calibration_mode='minmax', per_channel=True, learned_rounding=False")
        self.calibration_mode = calibration_mode
        self.smooth_transitions = smooth_transitions
        self.per_channel = per_channel
        self.momentum = momentum
        self.process_group = process_group
        self.sync_ranges = sync_ranges
        self.register_buffer("running_min", None)
        self.register_buffer("running_max", None)
        self.register_buffer("num_batches_tracked", torch.tensor(0))
        self.register_buffer("stats_frozen", torch.tensor(False))
        self.learned_rounding = None
        self.soft_mask = LearnedSoftMask() if smooth_transitions else None
        self.register_buffer("calibration_histogram", None)
        self.histogram_bins = 2048
        self._frozen_py = None      # host copy of stats_frozen: no device sync on the hot path
        self._packed_cache = None   # (tensor, version, packed ranges) of the most recent sweep
        # feature-level distillation (train.py:599-610), optional: set `kd_teacher` to the teacher's
        # fp32 feature map of this layer before the student's forward; the training forward then
        # also leaves mse(features_q, teacher) in `kd_feature_loss` (differentiable)
        self.kd_teacher = None
        self.kd_feature_loss = None

    # -- state ---------------------------------------------------------------------------------
    def _load_from_state_dict(self, state_dict, prefix, local_metadata, strict, missing_keys,
                              unexpected_keys, error_msgs):
        for name in ("running_min", "running_max"):
            key = prefix + name
            if key in state_dict and getattr(self, name) is None:
                setattr(self, name, torch.zeros_like(state_dict[key]))
        super()._load_from_state_dict(state_dict, prefix, local_metadata, strict, missing_keys,
                                      unexpected_keys, error_msgs)
        self._frozen_py = None

    def _is_frozen(self) -> bool:
        if self._frozen_py is None:
            self._frozen_py = bool(self.stats_frozen)      # one sync, then cached on the host
        return self._frozen_py

    def freeze_calibration(self):
        self.stats_frozen = torch.tensor(True, device=self.stats_frozen.device)
        self._frozen_py = True

    # -- ranges --------------------------------------------------------------------------------
    def _batch_ranges(self, x: torch.Tensor) -> torch.Tensor:
        """packed [min, -max] of this batch (all ranks when sync_ranges is set); decoded once per sweep."""
        hit = self._packed_cache
        if hit is not None and hit[0] is x and hit[1] == x._version:
            return hit[2]
        _, _, keys = channel_sweep(x, want_ranges=True)
        packed = ops.ranges_decode(keys)
        if self.sync_ranges:
            allreduce_ranges(packed, self.process_group)
        self._packed_cache = (x, x._version, packed)
        return packed

    @torch.no_grad()
    def update_running_stats(self, x: torch.Tensor):
        """EMA of per-channel min/max (quantization.py:319-353)."""
        if self._is_frozen():
            return
        C = x.shape[1]
        first = self.running_min is None
        if first:
            self.running_min = torch.empty((1, C, 1, 1), device=x.device, dtype=torch.float32)
            self.running_max = torch.empty((1, C, 1, 1), device=x.device, dtype=torch.float32)
        import torch.distributed as dist
        if self.sync_ranges and dist.is_available() and dist.is_initialized() and dist.get_world_size(self.process_group) > 1:
            packed = self._batch_ranges(x)                   # decode, all-reduce over the ranks, then the EMA
            ops.ranges_ema(packed, self.running_min, self.running_max, self.momentum, first)
        else:
            # K1's epilogue in one launch: decode + EMA; the batch ranges stay cached for the calibration pass,
            # which quantises with them (quantization.py:415-417)
            _, _, keys = channel_sweep(x, want_ranges=True)
            packed = ops.ranges_finish(keys, self.running_min, self.running_max, self.momentum, first)
            self._packed_cache = (x, x._version, packed)
        self.num_batches_tracked += 1

    def _qtable(self, x: torch.Tensor, training: bool) -> torch.Tensor:
        # the reference decides on the MODULE flag, not on the call argument (quantization.py:415-417):
        # in MCAQYOLO.calibrate() the model is in eval mode and the hook passes training=True, so the EMA
        # is updated but the calibration pass itself quantises with the current batch's min / max
        use_running = self.running_min is not None and (self.training or self._is_frozen())
        if use_running:
            return ops.build_qtable(None, self.running_min, self.running_max)
        return ops.build_qtable(self._batch_ranges(x))

    # -- forward -------------------------------------------------------------------------------
    def forward(self, x: torch.Tensor, bit_map: torch.Tensor, training: Optional[bool] = None) -> torch.Tensor:
        if training is None:
            training = self.training
        if x.dim() != 4:
            raise ValueError("expected (B, C, H, W)")
        assert x.shape[0] == bit_map.shape[0], f"Batch size mismatch: {x.shape[0]} vs {bit_map.shape[0]}"
        if training:
            self.update_running_stats(x)
        qtable = self._qtable(x, training)
        mask = None
        if self.smooth_transitions and self.soft_mask is not None:
            mask = self.soft_mask(bit_map, x)
        needs_grad = torch.is_grad_enabled() and (
            x.requires_grad or bit_map.requires_grad or (mask is not None and mask.requires_grad))
        self.kd_feature_loss = None
        if training or needs_grad:
            # fractional-bit compose; with an integer map it reduces to per-tile STE quantisation
            xc = x if x.is_contiguous() else x.contiguous()
            t = self.kd_teacher
            if t is not None and t.shape == x.shape:
                if ops.kd_geometry_ok(xc, bit_map):
                    y, self.kd_feature_loss = _FractionalQuantKD.apply(xc, bit_map.float(), mask, qtable, t.detach())
                else:       # geometry outside the vector kernels: same loss, composed
                    y = _FractionalQuant.apply(xc, bit_map.float(), mask, qtable)
                    self.kd_feature_loss = F.mse_loss(y.float(), t.detach().float())
            else:
                y = _FractionalQuant.apply(xc, bit_map.float(), mask, qtable)
        else:
            m3 = None if mask is None else mask.reshape(mask.shape[0], mask.shape[-2], mask.shape[-1])
            y = ops.tile_quantize(x, bit_map, qtable, m3)
        _SWEEP.clear()
        self._packed_cache = None
        return y

    def extra_repr(self) -> str:
        return (f"calibration_mode={self.calibration_mode}, smooth_transitions={self.smooth_transitions}, "
                f"per_channel={self.per_channel}")


# ---------------------------------------------------------------------------------------------
# hook composition and drop-in helpers
# ---------------------------------------------------------------------------------------------
def mcaq_hook_forward(feat: torch.Tensor, analyzer, mapper, quantizer, temperature: float = 1.0,
                      quantize: bool = True, training: bool = False, calibrating: bool = False,
                      normalize_complexity: bool = False, layer: int = -1) -> dict:
    """Body of the reference's forward hook (models/mcaq_yolo.py:409-455) for one scale: returns
    the aux record {'layer','complexity','bit_map','features_q'}."""
    complexity = analyzer(feat)
    if normalize_complexity:
        B = complexity.shape[0]
        flat = complexity.reshape(B, -1)
        lo = torch.quantile(flat, 0.02, dim=1, keepdim=True).unsqueeze(-1)
        hi = torch.quantile(flat, 0.98, dim=1, keepdim=True).unsqueeze(-1)
        complexity = ((complexity - lo) / (hi - lo + 1e-8)).clamp(0.0, 1.0)
    bit_map = mapper(complexity, temperature, return_continuous=training)
    feat_q = quantizer(feat, bit_map, training=training or calibrating) if quantize else feat
    rec = {"layer": layer, "complexity": complexity, "bit_map": bit_map, "features_q": feat_q}
    kd = getattr(quantizer, "kd_feature_loss", None) if quantize else None
    if kd is not None:
        rec["kd_feature_loss"] = kd        # extra key: mse(features_q, quantizer.kd_teacher), fused in K3
    return rec


class _McaqCudaOpsShim:
    """Stand-in for the reference's extension module `mcaq_cuda_ops` (ops/src/mcaq_ops.cpp:70-77)."""
    __name__ = "mcaq_cuda_ops"
    spatial_quantize = staticmethod(ops.spatial_quantize)


def install_mcaq_cuda_ops():
    """Register the shim as `sys.modules['mcaq_cuda_ops']` so that the reference's
    `import mcaq_cuda_ops` (core/quantization.py:14-16) binds to this library."""
    import sys
    import types
    mod = types.ModuleType("mcaq_cuda_ops")
    mod.spatial_quantize = ops.spatial_quantize
    mod.__doc__ = "MCAQ spatial quantization (B200-native, libmcaq_b200.so)"
    sys.modules["mcaq_cuda_ops"] = mod
    return mod


def install(model, device=None, fused: bool = True):
    """Swap the three hot-path objects of a reference MCAQYOLO (or any object exposing
    `complexity_analyzer`, `bit_mapper`, `quantizers`) for the native ones, keeping their weights.
    With `fused` (default) the reference's hook closures (`model._mcaq_hooks`) are replaced by
    `fused.FusedMcaqHook`, same protocol, three launches per scale at inference."""
    dev = device or next(model.complexity_analyzer.parameters()).device
    a_old = model.complexity_analyzer
    a_new = MorphologicalComplexityAnalyzer(grid_size=a_old.grid_size, device=dev)
    a_new.load_state_dict(a_old.state_dict())
    a_new.train(a_old.training)
    model.complexity_analyzer = a_new
    m_old = model.bit_mapper
    if type(m_old).__name__ == "LinearBitMapper":
        m_new = LinearBitMapper(int(m_old.min_bits), int(m_old.max_bits), m_old.eps_spread)
    elif type(m_old).__name__ == "ComplexityToBitMappingNetwork":
        m_new = ComplexityToBitMappingNetwork(int(m_old.min_bits), int(m_old.max_bits),
                                              enforce_monotonicity=m_old.enforce_monotonicity).to(dev)
        m_new.load_state_dict(m_old.state_dict())
    else:
        m_new = m_old            # arbitrary user mapper (scripts/m3_permutation.py, m4_variation_gain.py)
    if m_new is not m_old:
        m_new.train(m_old.training)
        model.bit_mapper = m_new
    for key in list(model.quantizers.keys()):
        q_old = model.quantizers[key]
        q_new = SpatialAdaptiveQuantization(smooth_transitions=q_old.smooth_transitions,
                                            momentum=q_old.momentum).to(dev)
        q_new.load_state_dict(q_old.state_dict())
        q_new.train(q_old.training)
        model.quantizers[key] = q_new
    hooks = getattr(model, "_mcaq_hooks", None)
    if fused and hooks is not None and hasattr(model, "backbone_out_indices"):
        from .fused import FusedMcaqHook
        for h in hooks:
            h.remove()
        layers = list(model.model.model)
        model._mcaq_hooks = [layers[idx].register_forward_hook(FusedMcaqHook(model, idx))
                             for idx in model.backbone_out_indices if 0 <= idx < len(layers)]
    return model


def build_fixture_modules(W: dict, device="cuda", grid_size: int = 8, linear_mapper: bool = False):
    """Modules loaded from dict-of-numpy state_dicts (tests/golden/weights.npz layout)."""
    def sd(d):
        return {k: torch.as_tensor(v) for k, v in d.items()}
    a = MorphologicalComplexityAnalyzer(grid_size=grid_size, device=device)
    a.load_state_dict(sd(W["analyzer"]))
    if linear_mapper:
        m = LinearBitMapper()
    else:
        m = ComplexityToBitMappingNetwork().to(device)
        m.load_state_dict(sd(W["mapper"]))
    q = SpatialAdaptiveQuantization().to(device)
    q.load_state_dict(sd(W["quantizer"]))
    return a.eval(), m.eval(), q.eval()
