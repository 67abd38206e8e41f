"""Tensor-level wrappers over the C ABI (include/mcaq_b200.h).

PyTorch is plumbing here: it owns device memory and the current stream; every function below
validates its arguments, allocates outputs with the caching allocator and forwards raw
pointers to libmcaq_b200.so on `torch.cuda.current_stream()`.  Nothing synchronises, so all
of it is CUDA-graph capturable.  No CPU fallback: non-CUDA tensors raise.
"""
from __future__ import annotations

import torch

from . import _lib
from ._lib import MCAQ_BF16, MCAQ_F16, MCAQ_F32, check


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _dtype_code(t: torch.Tensor) -> int:
    if t.dtype == torch.float32:
        return MCAQ_F32
    if t.dtype == torch.bfloat16:
        return MCAQ_BF16
    if t.dtype == torch.float16:          # hooked backbone outputs under autocast (train.py:192, 582, 748)
        return MCAQ_F16
    raise TypeError(f"mcaq_b200 supports float32 / bfloat16 / float16 feature maps, got {t.dtype}")


def _need_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise RuntimeError("mcaq_b200 ops need CUDA tensors (there is no CPU fallback in this package)")


def _f32c(t: torch.Tensor) -> torch.Tensor:
    t = t if t.dtype == torch.float32 else t.float()
    return t if t.is_contiguous() else t.contiguous()


def _ptr(t):
    return None if t is None else t.data_ptr()


# Optional per-launch hook `f(name, phase)` (phase 0 before, 1 after the launch); bench.py sets
# it to bracket kernels with CUDA events for the roofline attribution pass.  LAUNCHES counts
# kernel launches issued through this module.
EVENT_HOOK = None
LAUNCHES = 0
# K3 input staging: False = LDG.128 into registers (tile_quantize_vec_kernel), True = cp.async.bulk + mbarrier into shared
# memory (tile_quantize_tma_kernel); both bit-identical (tests/test_gpu_round2.py).  Default from the measurements in
# profiles/r02_k3_tma.txt; MCAQ_K3_TMA=0/1 overrides.
import os as _os
K3_TMA = bool(int(_os.environ.get("MCAQ_K3_TMA", "0")))


def _call(name: str, *args):
    global LAUNCHES
    fn = getattr(_lib.load(), name)
    LAUNCHES += 1
    if EVENT_HOOK is None:
        rc = fn(*args)
    else:
        EVENT_HOOK(name, 0)
        rc = fn(*args)
        EVENT_HOOK(name, 1)
    if rc is not None:
        check(rc, name)


def is_nhwc(x: torch.Tensor) -> bool:
    """True when x is a channels_last tensor the NHWC kernels cover (memory (B,H,W,C) dense, C a
    multiple of 16 with C / (16 / itemsize) a power of two <= 256, 16-byte aligned)."""
    if x.dim() != 4 or x.is_contiguous() or not x.is_contiguous(memory_format=torch.channels_last):
        return False
    if x.dtype == torch.float16:          # NHWC kernels: fp32 / bf16 (fp16 channels_last is copied to NCHW)
        return False
    C = x.shape[1]
    vec = 16 // x.element_size()
    vpp = C // vec
    return C % 16 == 0 and vpp >= 1 and vpp <= 256 and (vpp & (vpp - 1)) == 0 and x.data_ptr() % 16 == 0


def as_kernel_layout(x: torch.Tensor) -> torch.Tensor:
    """x unchanged if it is NCHW-contiguous or a covered channels_last tensor, else an NCHW copy."""
    return x if (x.is_contiguous() or is_nhwc(x)) else x.contiguous()


def tile_size(H: int, grid_size: int) -> int:
    """morphology.py:359-376."""
    return _lib.load().mcaq_tile_size(int(H), int(grid_size))


# ----------------------------------------------------------------------------- K1
def reduce_planes(x: torch.Tensor, want_ranges: bool = True):
    """One sweep of x (B,C,H,W): returns (sum_c x, sum_c |x|, keys) with keys the int32 (2C,)
    ordered-int min/max accumulators (None when want_ranges is False)."""
    _need_cuda(x)
    if x.dim() != 4:
        raise ValueError("expected (B,C,H,W)")
    x = as_kernel_layout(x)
    B, C, H, W = x.shape
    s = torch.empty((B, H, W), device=x.device, dtype=torch.float32)
    a = torch.empty((B, H, W), device=x.device, dtype=torch.float32)
    keys = None
    if want_ranges:
        keys = torch.empty((2 * C,), device=x.device, dtype=torch.int32)
        _call("mcaq_ranges_reset", keys.data_ptr(), C, _stream())
    reduce_planes_into(x, s, a, keys)
    return s, a, keys


def reduce_planes_into(x: torch.Tensor, s: torch.Tensor, a: torch.Tensor, keys: torch.Tensor | None):
    """K1 into caller-owned planes / armed keys; x NCHW-contiguous or covered channels_last."""
    B, C, H, W = x.shape
    name = "mcaq_reduce_planes" if x.is_contiguous() else "mcaq_reduce_planes_nhwc"
    _call(name, x.data_ptr(), _dtype_code(x), B, C, H, W, s.data_ptr(), a.data_ptr(), _ptr(keys), _stream())


def ranges_decode(keys: torch.Tensor) -> torch.Tensor:
    """keys -> packed (2C,) fp32 = [min, -max] (a single MIN all-reduce merges ranks)."""
    C = keys.numel() // 2
    packed = torch.empty((2 * C,), device=keys.device, dtype=torch.float32)
    _call("mcaq_ranges_decode", keys.data_ptr(), C, packed.data_ptr(), _stream())
    return packed


def ranges_ema(packed: torch.Tensor, running_min: torch.Tensor, running_max: torch.Tensor,
               momentum: float, first: bool):
    """In-place EMA of running_min/max (quantization.py:340-347)."""
    C = packed.numel() // 2
    assert running_min.numel() == C and running_max.numel() == C
    assert running_min.is_contiguous() and running_max.is_contiguous()
    _call("mcaq_ranges_ema", packed.data_ptr(), C, float(momentum), int(bool(first)),
                                      running_min.data_ptr(), running_max.data_ptr(), _stream())


def ranges_finish(keys: torch.Tensor, running_min: torch.Tensor, running_max: torch.Tensor, momentum: float,
                  first: bool) -> torch.Tensor:
    """keys -> packed [min, -max] AND the in-place EMA of running_min / running_max, one launch
    (quantization.py:319-353); returns packed (the batch ranges, which the calibration pass also quantises with)."""
    C = keys.numel() // 2
    assert running_min.numel() == C and running_max.numel() == C
    packed = torch.empty((2 * C,), device=keys.device, dtype=torch.float32)
    _call("mcaq_ranges_finish", keys.data_ptr(), C, float(momentum), int(bool(first)), running_min.data_ptr(),
          running_max.data_ptr(), packed.data_ptr(), _stream())
    return packed


def build_qtable(packed: torch.Tensor | None = None, running_min: torch.Tensor | None = None,
                 running_max: torch.Tensor | None = None) -> torch.Tensor:
    """(7, C, 2) table of {scale, zero_point} for bits 2..8 (quantization.py:41-66)."""
    if packed is not None:
        C = packed.numel() // 2
        dev = packed.device
    else:
        running_min, running_max = _f32c(running_min).reshape(-1), _f32c(running_max).reshape(-1)
        C = running_min.numel()
        dev = running_min.device
    qt = torch.empty((7, C, 2), device=dev, dtype=torch.float32)
    _call("mcaq_build_qtable", _ptr(packed), _ptr(running_min), _ptr(running_max), C, qt.data_ptr(),
                                        _stream())
    return qt


# ----------------------------------------------------------------------------- K3
def _bitmap_args(x, bit_map):
    B, C, H, W = x.shape
    if bit_map.dim() != 3 or bit_map.shape[0] != B:
        raise RuntimeError(f"bit_map must be (B,Ht,Wt) with B={B}, got {tuple(bit_map.shape)}")
    return B, C, H, W, int(bit_map.shape[1]), int(bit_map.shape[2])


def tile_quantize(x: torch.Tensor, bit_map: torch.Tensor, qtable: torch.Tensor,
                  mask: torch.Tensor | None = None, out: torch.Tensor | None = None,
                  want_codes: bool = False):
    """y = m * Q_{b_T(p)}(x) (Eq.19; quantization.py:729-744).  `out` may be x (in place)."""
    _need_cuda(x, bit_map, qtable, mask)
    x = x if x.is_contiguous() else x.contiguous()
    B, C, H, W, Ht, Wt = _bitmap_args(x, bit_map)
    bit_map = _f32c(bit_map)
    if qtable.shape != (7, C, 2):
        raise RuntimeError(f"qtable must be (7,{C},2)")
    if mask is not None:
        mask = _f32c(mask)
        if mask.numel() != B * H * W:
            raise RuntimeError("mask must be (N, 1, H, W)")
    y = torch.empty_like(x) if out is None else out
    codes = torch.empty(x.shape, device=x.device, dtype=torch.int8) if want_codes else None
    _call("mcaq_tile_quantize", x.data_ptr(), y.data_ptr(), _dtype_code(x), B, C, H, W,
                                         bit_map.data_ptr(), Ht, Wt, qtable.data_ptr(), _ptr(mask),
                                         _ptr(codes), _stream())
    return (y, codes) if want_codes else y


def tile_quantize_train_fwd(x, bit_map, qtable, mask=None):
    _need_cuda(x, bit_map, qtable, mask)
    x = x if x.is_contiguous() else x.contiguous()
    B, C, H, W, Ht, Wt = _bitmap_args(x, bit_map)
    bit_map = _f32c(bit_map)
    mask = None if mask is None else _f32c(mask)
    y = torch.empty_like(x)
    _call("mcaq_tile_quantize_train_fwd", x.data_ptr(), y.data_ptr(), _dtype_code(x), B, C, H, W,
                                                   bit_map.data_ptr(), Ht, Wt, qtable.data_ptr(), _ptr(mask),
                                                   _stream())
    return y


def kd_geometry_ok(x: torch.Tensor, bit_map: torch.Tensor) -> bool:
    """Whether the distillation-fused entry points cover this geometry (the vector path of K3:
    every aligned group of 4 pixels inside one row and one tile; every YOLO feature map)."""
    H, W, Wt = int(x.shape[2]), int(x.shape[3]), int(bit_map.shape[-1])
    vec = 16 // x.element_size()
    return (H * W) % vec == 0 and W % 4 == 0 and W % Wt == 0 and (W // Wt) % 4 == 0


def _teacher_arg(teacher, x):
    _need_cuda(teacher)
    if teacher.shape != x.shape:
        raise RuntimeError(f"teacher must have the shape of x {tuple(x.shape)}, got {tuple(teacher.shape)}")
    return _f32c(teacher)


def tile_quantize_train_fwd_kd(x, bit_map, qtable, mask, teacher):
    """Training forward with the feature-distillation term (train.py:599-610) folded in: returns
    (y, sum((y - teacher)^2) as a 0-dim fp64 tensor).  mse = sum / y.numel()."""
    _need_cuda(x, bit_map, qtable, mask)
    x = x if x.is_contiguous() else x.contiguous()
    B, C, H, W, Ht, Wt = _bitmap_args(x, bit_map)
    bit_map = _f32c(bit_map)
    mask = None if mask is None else _f32c(mask)
    teacher = _teacher_arg(teacher, x)
    y = torch.empty_like(x)
    kd_sum = torch.zeros((), device=x.device, dtype=torch.float64)
    _call("mcaq_tile_quantize_train_fwd_kd", x.data_ptr(), y.data_ptr(), _dtype_code(x), B, C, H, W,
          bit_map.data_ptr(), Ht, Wt, qtable.data_ptr(), _ptr(mask), teacher.data_ptr(), kd_sum.data_ptr(),
          _stream())
    return y, kd_sum


def _bwd_outputs(x, bit_map, mask):
    B, C, H, W, Ht, Wt = _bitmap_args(x, bit_map)
    dx = torch.empty_like(x)
    dbit = torch.zeros((B, Ht, Wt), device=x.device, dtype=torch.float32)
    dmask = torch.zeros((B, H, W), device=x.device, dtype=torch.float32) if mask is not None else None
    return dx, dbit, dmask


def tile_quantize_train_bwd(grad_y, x, bit_map, qtable, mask=None):
    """Returns (dx, dbit (B,Ht,Wt) fp32, dmask (B,H,W) fp32 or None)."""
    _need_cuda(grad_y, x, bit_map, qtable, mask)
    x = x if x.is_contiguous() else x.contiguous()
    grad_y = grad_y if grad_y.is_contiguous() else grad_y.contiguous()
    if grad_y.dtype != x.dtype:
        grad_y = grad_y.to(x.dtype)
    B, C, H, W, Ht, Wt = _bitmap_args(x, bit_map)
    bit_map = _f32c(bit_map)
    mask = None if mask is None else _f32c(mask)
    dx, dbit, dmask = _bwd_outputs(x, bit_map, mask)
    _call("mcaq_tile_quantize_train_bwd", grad_y.data_ptr(), x.data_ptr(), dx.data_ptr(),
                                                   _dtype_code(x), B, C, H, W, bit_map.data_ptr(), Ht, Wt,
                                                   qtable.data_ptr(), _ptr(mask), dbit.data_ptr(), _ptr(dmask),
                                                   _stream())
    return dx, dbit, dmask


def tile_quantize_train_bwd_kd(grad_y, x, bit_map, qtable, mask, teacher, kd_coef):
    """Backward of (y, mse) jointly: g_t = grad_y + kd_coef * (y - teacher) with y recomputed from x.
    kd_coef: 0-dim / 1-element fp32 CUDA tensor = dL/d(mse) * 2 / numel (stays on the device)."""
    _need_cuda(grad_y, x, bit_map, qtable, mask, kd_coef)
    x = x if x.is_contiguous() else x.contiguous()
    grad_y = grad_y if grad_y.is_contiguous() else grad_y.contiguous()
    if grad_y.dtype != x.dtype:
        grad_y = grad_y.to(x.dtype)
    B, C, H, W, Ht, Wt = _bitmap_args(x, bit_map)
    bit_map = _f32c(bit_map)
    mask = None if mask is None else _f32c(mask)
    teacher = _teacher_arg(teacher, x)
    kd_coef = _f32c(kd_coef.reshape(1))
    dx, dbit, dmask = _bwd_outputs(x, bit_map, mask)
    _call("mcaq_tile_quantize_train_bwd_kd", grad_y.data_ptr(), x.data_ptr(), dx.data_ptr(), _dtype_code(x),
          B, C, H, W, bit_map.data_ptr(), Ht, Wt, qtable.data_ptr(), _ptr(mask), teacher.data_ptr(),
          kd_coef.data_ptr(), dbit.data_ptr(), _ptr(dmask), _stream())
    return dx, dbit, dmask


def spatial_quantize(input: torch.Tensor, bit_map: torch.Tensor, min_vals: torch.Tensor,
                     max_vals: torch.Tensor, tile_h: int, tile_w: int,
                     mask: torch.Tensor | None = None) -> torch.Tensor:
    """The reference's extension op `mcaq_cuda_ops.spatial_quantize` (ops/src/mcaq_ops.cpp:22-77):
    same arguments, same checks (RuntimeError on shape mismatch), new fp32 output allocation."""
    _need_cuda(input, bit_map, min_vals, max_vals, mask)
    if input.dim() != 4:
        raise RuntimeError("input must be (N, C, H, W)")
    if input.dtype != torch.float32:
        raise RuntimeError("expected scalar type Float")      # data_ptr<float>() in the reference
    N, C, H, W = input.shape
    if min_vals.numel() != C or max_vals.numel() != C:
        raise RuntimeError(f"min_vals/max_vals must have one entry per channel (C={C}), got "
                           f"{min_vals.numel()} — expand per-tensor stats before calling")
    if mask is not None and mask.numel() != N * H * W:
        raise RuntimeError("mask must be (N, 1, H, W)")
    input = input.contiguous()
    bit_map = _f32c(bit_map)
    mn, mx = _f32c(min_vals).reshape(-1), _f32c(max_vals).reshape(-1)
    mask = None if mask is None else _f32c(mask)
    out = torch.empty_like(input)
    if bit_map.dim() != 3 or bit_map.shape[0] != N:
        raise RuntimeError(f"bit_map must be (N, Ht, Wt) with N={N}, got {tuple(bit_map.shape)}")
    # the reference's launcher symbol (void) + its per-thread status word: errors surface as RuntimeError
    lib = _lib.load()
    _call("launch_spatial_quantization", input.data_ptr(), bit_map.data_ptr(), mn.data_ptr(), mx.data_ptr(),
                                            _ptr(mask), out.data_ptr(), N, C, H, W, int(tile_h), int(tile_w),
                                            int(bit_map.shape[1]), int(bit_map.shape[2]), _stream())
    check(lib.mcaq_level0_status(), "launch_spatial_quantization")
    return out


# ----------------------------------------------------------------------------- K2
def morph_fits(B: int, C: int, H: int, W: int, grid_size: int) -> bool:
    """True when the fused per-image kernel covers the geometry (planes up to 160 columns, tiles up to 32)."""
    return bool(_lib.load().mcaq_morph_fits(int(B), int(C), int(H), int(W), int(grid_size)))


def morph_phi(sum_plane: torch.Tensor, C: int, grid_size: int, consts: torch.Tensor, debug: bool = False,
              force_planes: bool = False):
    """phi tiles (B,ht,wt,8) from the channel-sum plane (morphology.py:826-873).  Feature-map sized planes run
    in the fused per-image kernel; image-sized ones (640x640, 1280x1280: tiles of 64 / 128 pixels, the
    curriculum-scoring inputs of utils/dataset.py:345-353) in the plane pipeline (csrc/morph_planes.cu)."""
    _need_cuda(sum_plane, consts)
    B, H, W = sum_plane.shape
    tile = tile_size(H, grid_size)
    ht, wt = H // tile, W // tile
    if ht <= 0 or wt <= 0:
        raise RuntimeError(f"plane {H}x{W} is smaller than one {tile}-pixel tile")
    Hc, Wc = ht * tile, wt * tile
    dev = sum_plane.device
    sum_plane = _f32c(sum_plane)
    phi = torch.empty((B, ht, wt, 8), device=dev, dtype=torch.float32)
    dbg = {}
    fits = morph_fits(B, C, H, W, grid_size) and not force_planes      # force_planes: tests cross-check the two paths
    if debug:
        ww = (Wc + 31) // 32
        dbg = dict(gray=torch.empty((B, Hc, Wc), device=dev, dtype=torch.float32),
                   edge_bits=torch.zeros((B, Hc, ww), device=dev, dtype=torch.int32),
                   bin_bits=torch.zeros((B, Hc, ww), device=dev, dtype=torch.int32),
                   lbp_hist=torch.zeros((B, ht, wt, 10), device=dev, dtype=torch.int32),
                   counts=torch.zeros((B, ht, wt, 12), device=dev, dtype=torch.int32))
    if fits:
        _call("mcaq_morph_phi", sum_plane.data_ptr(), B, int(C), H, W, int(grid_size), consts.data_ptr(),
              phi.data_ptr(), _ptr(dbg.get("gray")), _ptr(dbg.get("edge_bits")),
              _ptr(dbg.get("bin_bits")), _ptr(dbg.get("lbp_hist")),
              _ptr(dbg.get("counts")), _stream())
    else:
        nbytes = _lib.load().mcaq_morph_planes_workspace(B, int(C), H, W, int(grid_size))
        if nbytes < 0:
            check(int(nbytes), "mcaq_morph_planes_workspace")
        ws = torch.empty((int(nbytes),), device=dev, dtype=torch.uint8)
        _call("mcaq_morph_phi_planes", sum_plane.data_ptr(), B, int(C), H, W, int(grid_size), ws.data_ptr(),
              int(nbytes), phi.data_ptr(), _ptr(dbg.get("gray")), _ptr(dbg.get("edge_bits")),
              _ptr(dbg.get("bin_bits")), _ptr(dbg.get("lbp_hist")), _ptr(dbg.get("counts")), _stream())
    return (phi, dbg) if debug else phi


def complexity(phi: torch.Tensor, cmlp: torch.Tensor, consts: torch.Tensor, want_raw: bool = False):
    """phi -> complexity map (B,ht,wt): MLP, bilateral filter, clamp (morphology.py:959-968)."""
    _need_cuda(phi, cmlp, consts)
    B, ht, wt, _ = phi.shape
    phi = _f32c(phi)
    out = torch.empty((B, ht, wt), device=phi.device, dtype=torch.float32)
    raw = torch.empty_like(out) if want_raw else None
    _call("mcaq_complexity", phi.data_ptr(), B, ht, wt, cmlp.data_ptr(), consts.data_ptr(), _ptr(raw),
                                      out.data_ptr(), _stream())
    return (out, raw) if want_raw else out


def bit_mapper(cmap: torch.Tensor, mapper: torch.Tensor | None, temperature, continuous: bool,
               min_bits: float = 2.0, max_bits: float = 8.0, eps_spread: float = 1e-3) -> torch.Tensor:
    """Eval-mode bit mapper; mapper=None selects the linear (quantile) mapper."""
    _need_cuda(cmap, mapper)
    cmap = _f32c(cmap)
    B, ht, wt = cmap.shape
    out = torch.empty_like(cmap)
    use_t = temperature is not None
    t = max(float(temperature), 0.1) if use_t else 1.0
    _call("mcaq_bit_mapper", cmap.data_ptr(), B, ht, wt, _ptr(mapper), t, int(use_t), int(continuous),
                                      float(min_bits), float(max_bits), float(eps_spread), out.data_ptr(),
                                      _stream())
    return out


MAPPER_FLOATS = 4612
MAPPER_STEPS_FLOATS = 12


def mapper_steps(ext_block: torch.Tensor, t: float, use_t: bool, min_bits: float, max_bits: float) -> torch.Tensor:
    """Fill the step table behind the mapper block in `ext_block` (MAPPER_FLOATS + MAPPER_STEPS_FLOATS
    floats) in place: mcaq_mapper_steps, bisection with the mapper kernel itself."""
    _need_cuda(ext_block)
    if ext_block.numel() != MAPPER_FLOATS + MAPPER_STEPS_FLOATS or ext_block.dtype != torch.float32:
        raise RuntimeError("extended mapper block must hold MAPPER_FLOATS + MAPPER_STEPS_FLOATS fp32 values")
    _call("mcaq_mapper_steps", ext_block.data_ptr(), float(t), int(use_t), float(min_bits), float(max_bits),
          ext_block.data_ptr() + 4 * MAPPER_FLOATS, _stream())
    return ext_block


def soft_mask(bit_map: torch.Tensor, abs_plane: torch.Tensor, C: int, softmask: torch.Tensor,
              want_tiles: bool = False):
    """m (B,H,W) (quantization.py:213-239) from the bit map and the sum_c|x| plane."""
    _need_cuda(bit_map, abs_plane, softmask)
    bit_map = _f32c(bit_map)
    B, H, W = abs_plane.shape
    Ht, Wt = int(bit_map.shape[1]), int(bit_map.shape[2])
    m = torch.empty((B, H, W), device=abs_plane.device, dtype=torch.float32)
    mt = torch.empty((B, Ht, Wt), device=abs_plane.device, dtype=torch.float32) if want_tiles else None
    _call("mcaq_soft_mask", bit_map.data_ptr(), Ht, Wt, abs_plane.data_ptr(), B, int(C), H, W,
                                     softmask.data_ptr(), _ptr(mt), m.data_ptr(), _stream())
    return (m, mt) if want_tiles else m


# ----------------------------------------------------------------------------- fused path
def morph_fused(sum_plane: torch.Tensor, abs_plane: torch.Tensor | None, C: int, grid_size: int,
                cmlp: torch.Tensor, mapper: torch.Tensor | None, softmask: torch.Tensor | None,
                temperature, continuous: bool = False, keys: torch.Tensor | None = None,
                min_bits: float = 2.0, max_bits: float = 8.0, eps_spread: float = 1e-3,
                want_phi: bool = False, xchg=None):
    """K2 in one launch: phi -> complexity -> bit map -> soft mask (include/mcaq_b200.h
    mcaq_morph_fused).  mapper=None selects the linear mapper.  Returns a dict with
    complexity (B,ht,wt), bit_map (B,ht,wt), mask (B,H,W) or None, packed ranges or None, phi or None."""
    _need_cuda(sum_plane, abs_plane, cmlp, mapper, softmask, keys)
    B, H, W = sum_plane.shape
    tile = tile_size(H, grid_size)
    ht, wt = H // tile, W // tile
    dev = sum_plane.device
    cpx = torch.empty((B, ht, wt), device=dev, dtype=torch.float32)
    bits = torch.empty((B, ht, wt), device=dev, dtype=torch.float32)
    mask = torch.empty((B, H, W), device=dev, dtype=torch.float32) if softmask is not None else None
    phi = torch.empty((B, ht, wt, 8), device=dev, dtype=torch.float32) if want_phi else None
    packed = torch.empty((2 * C,), device=dev, dtype=torch.float32) if keys is not None else None
    use_t = temperature is not None
    t = max(float(temperature), 0.1) if use_t else 1.0
    # mapper kind: 1 linear (quantile) mapper, 0 MLP block, 2 MLP block + step table (constants.pack_mapping_steps)
    kind = 1 if mapper is None else (2 if mapper.numel() == MAPPER_FLOATS + MAPPER_STEPS_FLOATS else 0)
    args = (sum_plane.data_ptr(), _ptr(abs_plane), B, int(C), H, W, int(grid_size), _ptr(keys),
            _ptr(packed), cmlp.data_ptr(), _ptr(mapper), kind, _ptr(softmask), t, int(use_t),
            int(continuous), float(min_bits), float(max_bits), float(eps_spread), _ptr(phi), cpx.data_ptr(),
            bits.data_ptr(), _ptr(mask))
    if xchg is not None and xchg.world > 1:
        # multi-GPU: the first CTA also publishes this rank's ranges to every rank (peer.RangeExchange)
        import ctypes
        _call("mcaq_morph_fused_xchg", *args, ctypes.addressof(xchg.peers), xchg.rank, xchg.world, _stream())
    else:
        _call("mcaq_morph_fused", *args, _stream())
    return {"complexity": cpx, "bit_map": bits, "mask": mask, "packed": packed, "phi": phi}


def tile_quantize_ranges(x: torch.Tensor, bit_map: torch.Tensor, packed: torch.Tensor | None = None,
                         running_min: torch.Tensor | None = None, running_max: torch.Tensor | None = None,
                         mask: torch.Tensor | None = None, out: torch.Tensor | None = None, xchg=None):
    """K3 with the per-channel ranges given directly (no table kernel).  With `xchg` (a
    peer.RangeExchange of world > 1, host-driven use) a one-CTA kernel first merges the ranks'
    published vectors; the fused path passes the already merged `packed` of morph_fused instead."""
    _need_cuda(x, bit_map, packed, running_min, running_max, mask)
    x = as_kernel_layout(x)
    B, C, H, W, Ht, Wt = _bitmap_args(x, bit_map)
    bit_map = _f32c(bit_map)
    if packed is None:
        running_min, running_max = _f32c(running_min).reshape(-1), _f32c(running_max).reshape(-1)
    mask = None if mask is None else _f32c(mask)
    y = torch.empty_like(x) if out is None else out          # empty_like keeps channels_last
    if not x.is_contiguous():                                 # covered channels_last input: NHWC kernel
        if xchg is not None and xchg.world > 1:
            packed = xchg.merged(x.device)
        if y.stride() != x.stride():
            raise RuntimeError("channels_last input needs a channels_last output buffer")
        _call("mcaq_tile_quantize_ranges_nhwc", x.data_ptr(), y.data_ptr(), _dtype_code(x), B, C, H, W,
              bit_map.data_ptr(), Ht, Wt, _ptr(packed), _ptr(running_min), _ptr(running_max), _ptr(mask), _stream())
        return y
    vec = 4 if x.dtype == torch.float32 else 8
    needs_ws = not ((H * W) % vec == 0 and W % 4 == 0 and W % Wt == 0 and (W // Wt) % 4 == 0)
    if K3_TMA and not needs_ws and C % 32 == 0 and y.data_ptr() != x.data_ptr() and (xchg is None or xchg.world <= 1):
        _call("mcaq_tile_quantize_ranges_tma", x.data_ptr(), y.data_ptr(), _dtype_code(x), B, C, H, W, bit_map.data_ptr(),
              Ht, Wt, _ptr(packed), _ptr(running_min), _ptr(running_max), _ptr(mask), _stream())
        return y
    ws = torch.empty((7, C, 2), device=x.device, dtype=torch.float32) if needs_ws else None
    if xchg is not None and xchg.world > 1:
        pws = torch.empty((2 * C,), device=x.device, dtype=torch.float32)
        _call("mcaq_tile_quantize_xchg", x.data_ptr(), y.data_ptr(), _dtype_code(x), B, C, H, W, bit_map.data_ptr(),
              Ht, Wt, xchg.local, xchg.world, _ptr(ws), _ptr(pws), _ptr(mask), _stream())
        return y
    _call("mcaq_tile_quantize_ranges", x.data_ptr(), y.data_ptr(), _dtype_code(x), B, C, H, W, bit_map.data_ptr(),
          Ht, Wt, _ptr(packed), _ptr(running_min), _ptr(running_max), _ptr(ws), _ptr(mask), _stream())
    return y
