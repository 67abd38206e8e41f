"""Deterministic synthetic feature maps built from integer arithmetic only, so
the golden generator (build container) and the tests (any machine) see
bit-identical fp32 inputs without storing them.

kinds:
  noise  -- Irwin-Hall(4) of 16-bit uniforms, ~N(0.3, 2.3^2)   (SURVEY 8d `randn*2+0.3`)
  smooth -- exact bilinear x8 upsampling of a coarse integer grid + small noise
            (the survey's low-frequency variant: wider phi3/phi5 spread)
"""
import numpy as np


def feature_map(kind: str, B: int, C: int, H: int, W: int, seed: int = 0) -> np.ndarray:
    rng = np.random.Generator(np.random.PCG64(seed))
    if kind == "noise":
        u = rng.integers(0, 65536, size=(4, B, C, H, W), dtype=np.int64).sum(axis=0) - 131070
        return ((u + 4915).astype(np.float64) / 16384.0).astype(np.float32)
    if kind == "smooth":
        gh, gw = H // 8 + 2, W // 8 + 2
        g = rng.integers(-32768, 32768, size=(B, C, gh, gw), dtype=np.int64)
        ys = np.arange(H)
        xs = np.arange(W)
        y0, fy = ys // 8, ys % 8
        x0, fx = xs // 8, xs % 8
        top = g[:, :, y0][:, :, :, x0] * (8 - fx) + g[:, :, y0][:, :, :, x0 + 1] * fx
        bot = g[:, :, y0 + 1][:, :, :, x0] * (8 - fx) + g[:, :, y0 + 1][:, :, :, x0 + 1] * fx
        up = top * (8 - fy)[None, None, :, None] + bot * fy[None, None, :, None]      # *64
        n = rng.integers(-2048, 2049, size=(B, C, H, W), dtype=np.int64)
        # bias per channel so channel ranges differ
        bias = rng.integers(-8, 9, size=(1, C, 1, 1), dtype=np.int64) * 65536
        return ((up + n * 16 + bias).astype(np.float64) / (64.0 * 8192.0)).astype(np.float32)
    raise ValueError(kind)


def integer_bit_map(B: int, Ht: int, Wt: int, seed: int = 0) -> np.ndarray:
    rng = np.random.Generator(np.random.PCG64(seed + 1000))
    return rng.integers(2, 9, size=(B, Ht, Wt)).astype(np.float32)


def fractional_bit_map(B: int, Ht: int, Wt: int, seed: int = 0) -> np.ndarray:
    """Continuous bits in [2, 8] on a 1/64 grid (exact in fp32), with some exact
    integers and some 8.0 entries (the q_hi == q_lo corner, quantization.py:721-724)."""
    rng = np.random.Generator(np.random.PCG64(seed + 2000))
    b = rng.integers(128, 513, size=(B, Ht, Wt)).astype(np.float32) / np.float32(64.0)
    flat = b.reshape(-1)
    flat[::7] = np.float32(8.0)
    flat[3::11] = np.rint(flat[3::11])
    return flat.reshape(B, Ht, Wt)
