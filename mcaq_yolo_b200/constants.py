"""Host-side constant block and parameter packers for the morphology / tile-net kernels.

The fixed stencils are the reference's (morphology.py:485-488, 566-570, 342-347, 585-612),
stored as the exact fp32 values torch produces (hex literals) so that the kernels, the oracle
and the reference share them bit-for-bit (pinned by tests/test_oracle_golden.py and
tests/test_host_logic.py).
"""
from __future__ import annotations

import math

import numpy as np
import torch

CONSTS_FLOATS = 192
CMLP_FLOATS = 2884
MAPPER_FLOATS = 4612
SOFTMASK_FLOATS = 196

_CANNY_TAPS = ['0x1.be5f10p-5', '0x1.f41fd8p-3', '0x1.9c4868p-2', '0x1.f41fd8p-3', '0x1.be5f10p-5']
_ADAPT_TAPS = ['0x1.20c256p-7', '0x1.bcb868p-6', '0x1.0ab508p-4', '0x1.f2464cp-4', '0x1.6a7e1cp-3',
               '0x1.9ac20ap-3', '0x1.6a7e1cp-3', '0x1.f2464cp-4', '0x1.0ab508p-4', '0x1.bcb868p-6',
               '0x1.20c256p-7']


def _outer(taps):
    g1 = np.array([float.fromhex(h) for h in taps], dtype=np.float32)
    return (g1[None, :] * g1[:, None]).astype(np.float32)


def _round64(fn, x):
    return fn(np.asarray(x, dtype=np.float32).astype(np.float64)).astype(np.float32)


def constant_block() -> np.ndarray:
    """The MCAQ_CONSTS_FLOATS block read by mcaq_morph_phi / mcaq_complexity
    (offsets: csrc/morph_phi.cu K_*)."""
    c = np.zeros(CONSTS_FLOATS, dtype=np.float32)
    c[0:25] = _outer(_CANNY_TAPS).ravel()
    c[25:146] = _outer(_ADAPT_TAPS).ravel()
    ax = np.arange(5, dtype=np.float32) - np.float32(2)
    yy, xx = np.meshgrid(ax, ax, indexing="ij")
    c[146:171] = _round64(np.exp, -(yy * yy + xx * xx) / np.float32(2 * 2.0 ** 2)).ravel()
    c[171:176] = _round64(np.log, np.array([2, 4, 8, 16, 32], dtype=np.float32))
    c[176:181] = _round64(np.exp, np.float32(-0.1) * np.arange(5, dtype=np.float32))
    c[181] = np.float32(180.0 / math.pi)
    c[182] = np.float32(4.0 * math.pi)
    c[183] = np.float32(math.log2(10.0))
    c[184] = np.float32(2 * 0.1 ** 2)
    return c


_CONST_CACHE: dict = {}


def device_constants(device) -> torch.Tensor:
    key = str(device)
    if key not in _CONST_CACHE:
        _CONST_CACHE[key] = torch.from_numpy(constant_block()).to(device)
    return _CONST_CACHE[key]


def _flat(*tensors) -> torch.Tensor:
    return torch.cat([t.detach().reshape(-1).float() for t in tensors]).contiguous()


def _cached(owner, tag: str, tensors, build):
    """Re-pack only when a parameter/buffer changed (in-place updates bump `_version`,
    re-assignment changes the data pointer); keeps the hot path free of packing kernels."""
    key = tuple((t.data_ptr(), t._version) for t in tensors)
    cache = owner.__dict__.setdefault("_mcaq_pack_cache", {})
    hit = cache.get(tag)
    if hit is not None and hit[0] == key:
        return hit[1]
    out = build()
    cache[tag] = (key, out)
    return out


def pack_complexity_mlp(seq) -> torch.Tensor:
    """`complexity_mlp` Sequential(Linear(8,64), LN, ReLU, Linear(64,32), LN, ReLU, Linear(32,1), Sigmoid)."""
    ts = [seq[0].weight, seq[0].bias, seq[1].weight, seq[1].bias, seq[3].weight, seq[3].bias,
          seq[4].weight, seq[4].bias, seq[6].weight, seq[6].bias]

    def build():
        # layout the kernels stage into shared memory as is: first / second layer weights transposed
        # to [k][unit] (conflict-free per-lane reads), three pad floats (16-byte copies)
        out = _flat(seq[0].weight.t(), seq[0].bias, seq[1].weight, seq[1].bias, seq[3].weight.t(), seq[3].bias,
                    seq[4].weight, seq[4].bias, seq[6].weight, seq[6].bias, seq[6].bias.new_zeros(3))
        assert out.numel() == CMLP_FLOATS, out.numel()
        return out
    return _cached(seq, "cmlp", ts, build)


def fold_batchnorm(bn):
    """Eval BatchNorm1d as y = x*alpha + beta, the way ATen's CPU kernel forms it:
    invstd = 1/sqrt(var + eps) in fp64 -> fp32, alpha = invstd*gamma, beta = bias - mean*alpha."""
    invstd = (1.0 / torch.sqrt(bn.running_var.detach().double() + bn.eps)).float()
    alpha = invstd * bn.weight.detach().float()
    beta = bn.bias.detach().float() - bn.running_mean.detach().float() * alpha
    return alpha, beta


def pack_mapping_network(seq) -> torch.Tensor:
    """`mapping_network` Sequential(3x[Linear, BatchNorm1d, ReLU], Linear(32,1), Sigmoid), eval BN."""
    ts = []
    for li, bi in ((0, 1), (3, 4), (6, 7)):
        ts += [seq[li].weight, seq[li].bias, seq[bi].weight, seq[bi].bias, seq[bi].running_mean,
               seq[bi].running_var]
    ts += [seq[9].weight, seq[9].bias]

    def build():
        parts = []
        for li, bi in ((0, 1), (3, 4), (6, 7)):
            a, b = fold_batchnorm(seq[bi])
            parts += [seq[li].weight.t(), seq[li].bias, a, b]          # weights as [k][unit]
        parts += [seq[9].weight, seq[9].bias, seq[9].bias.new_zeros(3)]
        out = _flat(*parts)
        assert out.numel() == MAPPER_FLOATS, out.numel()
        return out
    return _cached(seq, "mapper", ts, build)


MAPPER_STEPS_FLOATS = 12


def mapping_is_monotone(mapper) -> bool:
    """True when bits(c) of the MLP mapper is provably non-decreasing in c, the condition under which its
    eval output may be read off a step table: Eq.18's constraint is declared (`enforce_monotonicity`) AND
    actually holds -- every Linear weight and every BatchNorm gamma (the sign of the folded eval scale)
    is >= 0.  A checkpoint saved before `enforce_weight_constraints()` ran, or a mapper built with
    enforce_monotonicity=False, fails this and is evaluated as the network it is.  One device->host read
    per weights version (cached like the packed blocks)."""
    if not getattr(mapper, "enforce_monotonicity", False):
        return False
    seq = mapper.mapping_network
    ts = [seq[i].weight for i in (0, 1, 3, 4, 6, 7, 9)]

    def build():
        return bool(torch.stack([t.detach().min() for t in ts]).min().item() >= 0.0)
    return _cached(seq, "monotone", ts, build)


def pack_mapping_steps(seq, temperature, min_bits: float, max_bits: float) -> torch.Tensor:
    """The mapper block followed by its step table (include/mcaq_b200.h: mcaq_mapper_steps) for this
    temperature and bit range: eval-mode integer bit maps of the fused kernel are read off the
    staircase of the (monotone) mapper instead of evaluating 4.2k FMAs per tile.  One tiny launch per
    (weights version, temperature, range), cached on the module like the block itself."""
    from . import ops
    base = pack_mapping_network(seq)
    use_t = temperature is not None
    t = max(float(temperature), 0.1) if use_t else 1.0
    cache = seq.__dict__.setdefault("_mcaq_steps_cache", {})
    if cache.get("base") is not base:
        cache.clear()
        cache["base"] = base
    key = (t, use_t, float(min_bits), float(max_bits))
    ext = cache.get(key)
    if ext is None:
        ext = torch.empty(MAPPER_FLOATS + MAPPER_STEPS_FLOATS, device=base.device, dtype=torch.float32)
        ext[:MAPPER_FLOATS].copy_(base)
        ops.mapper_steps(ext, t, use_t, min_bits, max_bits)
        cache[key] = ext
    return ext


def pack_soft_mask(soft_mask) -> torch.Tensor:
    """LearnedSoftMask: net[0] conv3x3(2->8), net[2] conv1x1(8->2), smooth_kernel (1,1,5,5)."""
    ts = [soft_mask.net[0].weight, soft_mask.net[0].bias, soft_mask.net[2].weight,
          soft_mask.net[2].bias, soft_mask.smooth_kernel]

    def build():
        out = _flat(*ts, soft_mask.smooth_kernel.new_zeros(1))
        assert out.numel() == SOFTMASK_FLOATS, out.numel()
        return out
    return _cached(soft_mask, "softmask", ts, build)
