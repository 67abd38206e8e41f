"""
CPU oracle for the MCAQ hot path (TEST INFRASTRUCTURE -- not product code).

This file is a numpy restatement of the reference's algorithm for the path
`complexity analyzer -> bit mapper -> soft mask -> tile-wise quantize`
(SURVEY.md section 8a).  Only `tests/`, `__graft_entry__.smoke()` and the
`cpu_baseline` / `--impl reference` legs of `bench.py` may import it; the
product path (`mcaq_yolo_b200/`) never does and fails loudly when the CUDA
library is missing.

Parity pin: every function below is checked against outputs of the real
reference (imported from /root/reference in the build container) stored
under `tests/golden/` by `tools/make_golden.py`.

Arithmetic contract (the same contract the CUDA kernels implement, so that
kernel == oracle bit-for-bit on every discrete quantity):

* everything is IEEE fp32; every numpy op below is one rounding.
* stencils (the reference's `F.conv2d`) are a chain of fused multiply-adds
  over the taps in row-major order starting from 0 -- measured to be
  bit-identical to torch 2.11 CPU conv2d for the 1-channel 3x3 / 5x5 / 11x11
  cases the path uses.
* channel sums (`x.mean(1)`, `x.abs().mean(1)`) use torch's CPU cascade
  order: 16-channel chunks summed sequentially, chunk totals accumulated
  sequentially, folded every 256 channels -- bit-identical to torch wherever
  torch's vectorised path handles the pixel (all pixels when H*W % 32 == 0).
* `nn.Linear` is a sequential FMA chain over k starting from 0 with the bias
  added last (bit-identical to torch CPU for out_features > 1).
* tile float sums are "column then row": each tile column top-to-bottom, then
  the column sums left-to-right (one GPU thread owns a column); LayerNorm
  statistics use a 32-lane butterfly tree.
* Otsu cumulative sums accumulate in fp64 and round to fp32 per element
  (what torch.cumsum does on CPU; exact, hence order independent).
* transcendental functions are evaluated in fp64 and rounded to fp32.

All reference citations are `file:line` under /root/reference/mcaq_yolo/.
"""

from __future__ import annotations

import math

import numpy as np

f32 = np.float32
f64 = np.float64


# ----------------------------------------------------------------------------
# exact fp32 primitives
# ----------------------------------------------------------------------------

def fma32(a, b, c):
    """Exact fp32 fused multiply-add, rn(a*b + c) with a single rounding.

    a*b is exact in fp64 (48 significant bits).  The fp64 sum may round; when
    that rounded sum sits exactly on an fp32 rounding tie the discarded error
    term decides the direction (sticky-bit correction), so the result equals
    C's fmaf() / CUDA's fmaf() for all finite inputs.
    """
    a, b, c = np.broadcast_arrays(np.asarray(a, dtype=f32), np.asarray(b, dtype=f32),
                                  np.asarray(c, dtype=f32))
    shape = a.shape
    a = a.astype(f64).reshape(-1)
    b = b.astype(f64).reshape(-1)
    c = c.astype(f64).reshape(-1)
    p = a * b                       # exact
    s = p + c                       # one fp64 rounding
    # TwoSum error term: s + e == p + c exactly
    bb = s - p
    e = (p - (s - bb)) + (c - bb)
    tie = (s.view(np.int64) & 0x1FFFFFFF) == 0x10000000
    fix = tie & (e != 0)
    if np.any(fix):
        toward = np.where(e > 0, np.inf, -np.inf)
        s[fix] = np.nextafter(s[fix], toward[fix])
    return s.astype(f32).reshape(shape)


def _t64(fn, x):
    """Transcendental in fp64, rounded once to fp32."""
    return fn(np.asarray(x, dtype=f32).astype(f64)).astype(f32)


def exp32(x):
    return _t64(np.exp, x)


def log32(x):
    return _t64(np.log, x)


def log2_32(x):
    return _t64(np.log2, x)


def log1p32(x):
    return _t64(np.log1p, x)


def sqrt32(x):
    return np.sqrt(np.asarray(x, dtype=f32))   # IEEE correctly rounded


def rint32(x):
    return np.rint(np.asarray(x, dtype=f32))   # half-to-even == torch.round


# ----------------------------------------------------------------------------
# fixed constants (shared bit-for-bit with the CUDA side, see
# mcaq_yolo_b200/constants.py which recomputes them the same way)
# ----------------------------------------------------------------------------

def gaussian_kernel2d(k: int, sigma: float) -> np.ndarray:
    """g1 = exp(-x^2 / (2 sigma^2)) / sum, g2 = outer(g1, g1), all fp32.
    morphology.py:485-488 (k=5, sigma=1), 566-570 (k=11, sigma=2.0),
    quantization.py:204-210 (k=5, sigma=5/3)."""
    x = np.arange(k, dtype=f32) - f32(k // 2)
    num = -(x * x)                                      # -(x**2)
    den = f32(2.0 * sigma * sigma)
    g1 = exp32(num / den)
    tot = f32(0)
    for v in g1:                                        # torch sum of <16 elems is sequential
        tot = f32(tot + v)
    g1 = (g1 / tot).astype(f32)
    return (g1[None, :] * g1[:, None]).astype(f32)


def _outer(taps_hex):
    g1 = np.array([float.fromhex(h) for h in taps_hex], dtype=f32)
    return (g1[None, :] * g1[:, None]).astype(f32)


# Normalised 1-D taps exactly as torch 2.11 produces them (torch's `g1.sum()` order is
# not the sequential one for 11 taps, so the values are pinned here and checked against
# tests/golden/weights.npz `const.*`; gaussian_kernel2d() reproduces the 5-tap sets).
CANNY_BLUR = _outer(['0x1.be5f10p-5', '0x1.f41fd8p-3', '0x1.9c4868p-2', '0x1.f41fd8p-3',
                     '0x1.be5f10p-5'])                       # morphology.py:485-488
ADAPT_BLUR = _outer(['0x1.20c256p-7', '0x1.bcb868p-6', '0x1.0ab508p-4', '0x1.f2464cp-4',
                     '0x1.6a7e1cp-3', '0x1.9ac20ap-3', '0x1.6a7e1cp-3', '0x1.f2464cp-4',
                     '0x1.0ab508p-4', '0x1.bcb868p-6', '0x1.20c256p-7'])   # morphology.py:566-570
SOBEL_X = np.array([[-1, 0, 1], [-2, 0, 2], [-1, 0, 1]], dtype=f32)
SOBEL_Y = np.array([[-1, -2, -1], [0, 0, 0], [1, 2, 1]], dtype=f32)
RAD2DEG = f32(180.0 / math.pi)
FOUR_PI = f32(4.0 * math.pi)
LOG2_10 = f32(math.log2(10.0))
LBP_OFFSETS = [(-1, -1), (-1, 0), (-1, 1), (0, 1), (1, 1), (1, 0), (1, -1), (0, -1)]


def bilateral_spatial(k: int = 5, sigma: float = 2.0) -> np.ndarray:
    """morphology.py:342-347."""
    c = np.arange(k, dtype=f32) - f32(k // 2)
    yy, xx = np.meshgrid(c, c, indexing="ij")
    return exp32(-(yy * yy + xx * xx) / f32(2 * sigma ** 2)).astype(f32)


BILATERAL_SPATIAL = bilateral_spatial()
BILATERAL_RANGE_DEN = f32(2 * 0.1 ** 2)


def fractal_tables(tile: int):
    """log(s) and exp(-0.1 i) for s = 2, 4, .., tile (morphology.py:585-612)."""
    scales = []
    s = 2
    while s <= tile:
        scales.append(s)
        s *= 2
    x = log32(np.array(scales, dtype=f32))
    w = exp32(f32(-0.1) * np.arange(len(scales), dtype=f32))
    return scales, x, w


# ----------------------------------------------------------------------------
# geometry
# ----------------------------------------------------------------------------

def tile_size(H: int, grid_size: int = 8) -> int:
    """morphology.py:359-376."""
    raw = max(4, H // grid_size)
    return 1 << (raw.bit_length() - 1)


def nearest_index(out_size: int, in_size: int) -> np.ndarray:
    """F.interpolate(mode='nearest') source index (quantization.py:236, 737):
    min(floor(dst * (float)in/out), in-1), scale in fp32."""
    scale = f32(in_size) / f32(out_size)
    idx = np.floor(np.arange(out_size, dtype=f32) * scale).astype(np.int64)
    return np.minimum(idx, in_size - 1)


def adaptive_windows(in_size: int, out_size: int):
    """adaptive_avg_pool2d windows (quantization.py:224)."""
    st = [(i * in_size) // out_size for i in range(out_size)]
    en = [-((-(i + 1) * in_size) // out_size) for i in range(out_size)]
    return st, en


# ----------------------------------------------------------------------------
# channel reductions  (K1)
# ----------------------------------------------------------------------------

def cascade_sum(get, n: int):
    """Sum get(0..n-1) in torch's CPU cascade order (ATen SumKernel
    multi_row_sum, level_step = 16)."""
    z = np.zeros_like(get(0))
    acc = [z.copy() for _ in range(4)]
    i = 0
    while i + 16 <= n:
        for _ in range(16):
            acc[0] = acc[0] + get(i)
            i += 1
        for j in range(1, 4):
            acc[j] = acc[j] + acc[j - 1]
            acc[j - 1] = z.copy()
            if (i & (15 << (4 * j))) != 0:
                break
    while i < n:
        acc[0] = acc[0] + get(i)
        i += 1
    for j in range(1, 4):
        acc[0] = acc[0] + acc[j]
    return acc[0]


def channel_sums(x: np.ndarray):
    """sum_c x and sum_c |x| per pixel, fp32 (B,H,W).
    Feeds `features.mean(dim=1)` (morphology.py:837) and
    `x.abs().mean(1)` (quantization.py:224)."""
    x = np.asarray(x, dtype=f32)
    C = x.shape[1]
    s = cascade_sum(lambda c: x[:, c], C)
    a = cascade_sum(lambda c: np.abs(x[:, c]), C)
    return s.astype(f32), a.astype(f32)


def channel_minmax(x: np.ndarray):
    """Per-channel min/max over (B,H,W) (quantization.py:423-426, 650-654)."""
    x = np.asarray(x, dtype=f32)
    return x.min(axis=(0, 2, 3)).astype(f32), x.max(axis=(0, 2, 3)).astype(f32)


def ema_update(run_min, run_max, mn, mx, momentum: float = 0.99):
    """quantization.py:340-347 (first call adopts the batch statistics)."""
    if run_min is None:
        return mn.copy(), mx.copy()
    m = f32(momentum)
    om = f32(1 - momentum)
    return (m * run_min + om * mn).astype(f32), (m * run_max + om * mx).astype(f32)


# ----------------------------------------------------------------------------
# stencils
# ----------------------------------------------------------------------------

def stencil(img: np.ndarray, k2d: np.ndarray, pad_mode: str) -> np.ndarray:
    """'same' cross-correlation of (B,H,W) with a KxK kernel: FMA chain over
    taps in row-major order from 0.  pad_mode 'zero' | 'edge'."""
    K = k2d.shape[0]
    p = K // 2
    B, H, W = img.shape
    if pad_mode == "zero":
        ip = np.pad(img, ((0, 0), (p, p), (p, p)))
    else:
        ip = np.pad(img, ((0, 0), (p, p), (p, p)), mode="edge")
    acc = np.zeros((B, H, W), dtype=f32)
    for ky in range(K):
        for kx in range(K):
            w = k2d[ky, kx]
            if w == 0:
                continue                       # fma(v, 0, acc) == acc
            acc = fma32(ip[:, ky:ky + H, kx:kx + W], w, acc)
    return acc


def sobel(img: np.ndarray):
    """morphology.py:385-395 (zero padding=1)."""
    return stencil(img, SOBEL_X, "zero"), stencil(img, SOBEL_Y, "zero")


# ----------------------------------------------------------------------------
# tile reductions
# ----------------------------------------------------------------------------

def tile_sum_f32(p: np.ndarray, tile: int) -> np.ndarray:
    """Float tile sums, column-then-row order: each tile column top-to-bottom, then the
    column sums left-to-right.  p: (B,Hc,Wc) -> (B,ht,wt)."""
    B, Hc, Wc = p.shape
    ht, wt = Hc // tile, Wc // tile
    v = p.reshape(B, ht, tile, wt, tile)
    cols = np.zeros((B, ht, wt, tile), dtype=f32)
    for y in range(tile):
        cols = cols + v[:, :, y, :, :]
    out = np.zeros((B, ht, wt), dtype=f32)
    for x in range(tile):
        out = out + cols[:, :, :, x]
    return out


def tile_count(p: np.ndarray, tile: int) -> np.ndarray:
    """Integer tile counts of a {0,1} plane -> int64 (B,ht,wt)."""
    B, Hc, Wc = p.shape
    ht, wt = Hc // tile, Wc // tile
    return p.reshape(B, ht, tile, wt, tile).astype(np.int64).sum(axis=(2, 4))


# ----------------------------------------------------------------------------
# morphology  (K2)
# ----------------------------------------------------------------------------

def normalize01(gray: np.ndarray) -> np.ndarray:
    """morphology.py:378-383.  gray: (B,Hc,Wc)."""
    mn = gray.min(axis=(1, 2), keepdims=True)
    mx = gray.max(axis=(1, 2), keepdims=True)
    return ((gray - mn) / ((mx - mn) + f32(1e-8))).astype(f32)


def otsu_threshold(b01: np.ndarray, bins: int = 256):
    """morphology.py:397-418.  Returns (thr (B,), argmax bin (B,))."""
    B = b01.shape[0]
    centers = ((np.arange(bins, dtype=f32) + f32(0.5)) / f32(bins)).astype(f32)
    thr = np.zeros(B, dtype=f32)
    arg = np.zeros(B, dtype=np.int64)
    for b in range(B):
        v = b01[b].ravel()
        ok = (v >= 0) & (v <= 1)                         # histc ignores out-of-range
        idx = (v[ok] * f32(bins)).astype(np.int64)       # (v-0)*bins/(1-0), exact
        idx[idx == bins] = bins - 1
        hist = np.bincount(idx, minlength=bins).astype(f32)
        tot = f32(max(float(hist.sum()), 1.0))
        p = (hist / tot).astype(f32)
        omega = np.cumsum(p.astype(f64)).astype(f32)
        pc = (p * centers).astype(f32)
        mu = np.cumsum(pc.astype(f64)).astype(f32)
        mu_t = mu[-1]
        num = (mu_t * omega - mu).astype(f32)
        num = (num * num).astype(f32)
        den = (omega * (f32(1.0) - omega)).astype(f32) + f32(1e-12)
        sigma_b = (num / den).astype(f32)
        a = int(np.argmax(sigma_b))                      # first maximum
        arg[b] = a
        thr[b] = centers[a]
    return thr, arg


def _shift_edge(p: np.ndarray, dy: int, dx: int) -> np.ndarray:
    """p[y+dy, x+dx] with replicate borders (morphology.py:433-436)."""
    B, H, W = p.shape
    pp = np.pad(p, ((0, 0), (1, 1), (1, 1)), mode="edge")
    return pp[:, 1 + dy:1 + dy + H, 1 + dx:1 + dx + W]


def canny_nms(mag, gx, gy):
    """morphology.py:426-449."""
    ang = (np.arctan2(gy.astype(f64), gx.astype(f64)).astype(f32) * RAD2DEG).astype(f32)
    ang = np.where(ang < 0, (ang + f32(180.0)).astype(f32), ang)
    bins = [
        ((ang < f32(22.5)) | (ang >= f32(157.5)), (0, 1), (0, -1)),
        ((ang >= f32(22.5)) & (ang < f32(67.5)), (-1, 1), (1, -1)),
        ((ang >= f32(67.5)) & (ang < f32(112.5)), (-1, 0), (1, 0)),
        ((ang >= f32(112.5)) & (ang < f32(157.5)), (-1, -1), (1, 1)),
    ]
    nms = np.zeros_like(mag)
    for sel, (dy1, dx1), (dy2, dx2) in bins:
        keep = (mag >= _shift_edge(mag, dy1, dx1)) & (mag >= _shift_edge(mag, dy2, dx2))
        nms = np.where(sel & keep, mag, nms)
    return nms, ang


def dilate3(e: np.ndarray) -> np.ndarray:
    """3x3 max-pool, stride 1, -inf padding, on a boolean plane."""
    B, H, W = e.shape
    pp = np.pad(e, ((0, 0), (1, 1), (1, 1)))
    out = np.zeros_like(e)
    for dy in range(3):
        for dx in range(3):
            out = out | pp[:, dy:dy + H, dx:dx + W]
    return out


def canny_edges(gray: np.ndarray, hysteresis_iters: int = 8, detail: dict | None = None):
    """morphology.py:457-509 (cv2compat).  gray: (B,Hc,Wc) in [0,1] -> bool."""
    b01 = stencil(gray, CANNY_BLUR, "zero")
    b255 = (b01 * f32(255.0)).astype(f32)
    thr, arg = otsu_threshold(b01)
    thr255 = (thr * f32(255.0)).astype(f32)[:, None, None]
    gx, gy = sobel(b255)
    mag = (np.abs(gx) + np.abs(gy)).astype(f32)
    nms, ang = canny_nms(mag, gx, gy)
    strong = nms > thr255
    weak = nms > (f32(0.5) * thr255).astype(f32)
    edge = strong.copy()
    for _ in range(max(1, hysteresis_iters)):
        edge = edge | (weak & dilate3(edge))
    if detail is not None:
        detail.update(b01=b01, otsu_bin=arg, thr255=thr255[:, 0, 0], mag=mag, nms=nms,
                      strong=strong, weak=weak, angle=ang)
    return edge


def adaptive_binarize(gray: np.ndarray, C: float = 2.0, detail: dict | None = None):
    """morphology.py:550-573."""
    g255 = (gray * f32(255.0)).astype(f32)
    local_mean = stencil(g255, ADAPT_BLUR, "edge")
    if detail is not None:
        detail.update(local_mean=local_mean)
    return g255 > (local_mean - f32(C)).astype(f32)


def fractal_dimension_tiles(edge: np.ndarray, tile: int, detail: dict | None = None):
    """morphology.py:575-621.  edge bool (B,Hc,Wc) -> Df (B,ht,wt) in [1,2]."""
    B, Hc, Wc = edge.shape
    ht, wt = Hc // tile, Wc // tile
    scales, x, w = fractal_tables(tile)
    S = len(scales)
    if S < 2:
        return np.ones((B, ht, wt), dtype=f32)
    counts = []
    for s in scales:
        occ = edge.reshape(B, Hc // s, s, Wc // s, s).any(axis=(2, 4))
        k = tile // s
        counts.append(occ.reshape(B, ht, k, wt, k).astype(np.int64).sum(axis=(2, 4)))
    if detail is not None:
        detail.update(box_counts=np.stack(counts, 0))
    y = [log32(c.astype(f32) + f32(1.0)) for c in counts]

    def seqsum(terms):
        acc = np.zeros_like(terms[0], dtype=f32) if np.ndim(terms[0]) else f32(0)
        for t in terms:
            acc = (acc + t).astype(f32) if np.ndim(acc) else f32(acc + t)
        return acc

    w_sum = seqsum([w[i] for i in range(S)])
    x_mean = f32(seqsum([f32(w[i] * x[i]) for i in range(S)]) / w_sum)
    y_mean = (seqsum([(w[i] * y[i]).astype(f32) for i in range(S)]) / w_sum).astype(f32)
    cov = seqsum([((f32(w[i] * f32(x[i] - x_mean))) * (y[i] - y_mean)).astype(f32) for i in range(S)])
    var = seqsum([f32(w[i] * f32(f32(x[i] - x_mean) * f32(x[i] - x_mean))) for i in range(S)])
    df = (-(cov / f32(var + f32(1e-12)))).astype(f32)
    return np.clip(df, f32(1.0), f32(2.0)).astype(f32)


_LBP_LUT = None


def lbp_lut() -> np.ndarray:
    """256 -> 10 uniform-LBP label table.  Bit i of the code is neighbour i of
    LBP_OFFSETS (morphology.py:634-646)."""
    global _LBP_LUT
    if _LBP_LUT is None:
        lut = np.zeros(256, dtype=np.int64)
        for code in range(256):
            bits = [(code >> i) & 1 for i in range(8)]
            trans = sum(abs(bits[i] - bits[i - 1]) for i in range(8))
            lut[code] = sum(bits) if trans <= 2 else 9
        _LBP_LUT = lut
    return _LBP_LUT


def lbp_labels(gray: np.ndarray) -> np.ndarray:
    B, H, W = gray.shape
    gp = np.pad(gray, ((0, 0), (1, 1), (1, 1)), mode="edge")
    code = np.zeros((B, H, W), dtype=np.int64)
    for i, (dy, dx) in enumerate(LBP_OFFSETS):
        nb = gp[:, 1 + dy:1 + dy + H, 1 + dx:1 + dx + W]
        code |= (nb >= gray).astype(np.int64) << i
    return lbp_lut()[code]


def lbp_entropy_tiles(gray: np.ndarray, tile: int, detail: dict | None = None):
    """morphology.py:623-652."""
    B, Hc, Wc = gray.shape
    ht, wt = Hc // tile, Wc // tile
    lab = lbp_labels(gray)
    hist = np.zeros((B, 10, ht, wt), dtype=np.int64)
    for k in range(10):
        hist[:, k] = tile_count(lab == k, tile)
    if detail is not None:
        detail.update(lbp_label=lab, lbp_hist=hist)
    p = (hist.astype(f32) / f32(tile * tile)).astype(f32)
    ent = np.zeros((B, ht, wt), dtype=f32)
    for k in range(10):
        ent = (ent + (p[:, k] * log2_32(p[:, k] + f32(1e-10))).astype(f32)).astype(f32)
    ent = -ent
    return (ent / LOG2_10).astype(f32)


def gradient_variance_tiles(gx, gy, tile: int):
    """morphology.py:654-670."""
    n = f32(tile * tile)

    def tile_var(t):
        m = (tile_sum_f32(t, tile) / n).astype(f32)
        m2 = (tile_sum_f32((t * t).astype(f32), tile) / n).astype(f32)
        return np.maximum((m2 - (m * m).astype(f32)).astype(f32), f32(0.0))

    v = (tile_var(gx) + tile_var(gy)).astype(f32)
    return (v / (v + f32(1.0))).astype(f32)


def euler_x4_tiles(m: np.ndarray, tile: int) -> np.ndarray:
    """4 * Euler mass per tile as exact integers (morphology.py:672-706)."""
    B, Hc, Wc = m.shape
    mp = np.pad(m.astype(np.int64), ((0, 0), (1, 1), (1, 1)))
    idx = mp[:, :-1, :-1] + 2 * mp[:, :-1, 1:] + 4 * mp[:, 1:, :-1] + 8 * mp[:, 1:, 1:]
    q1 = np.isin(idx, [1, 2, 4, 8]).astype(np.int64)
    q3 = np.isin(idx, [7, 11, 13, 14]).astype(np.int64)
    qd = np.isin(idx, [6, 9]).astype(np.int64)
    e4 = (q1 - q3 - 2 * qd)[:, :Hc, :Wc]
    ht, wt = Hc // tile, Wc // tile
    return e4.reshape(B, ht, tile, wt, tile).sum(axis=(2, 4))


def erode3(m: np.ndarray) -> np.ndarray:
    """-max_pool(-m, 3, 1, 1): out-of-image neighbours are ignored."""
    B, H, W = m.shape
    pp = np.pad(m, ((0, 0), (1, 1), (1, 1)), constant_values=True)
    out = np.ones_like(m)
    for dy in range(3):
        for dx in range(3):
            out = out & pp[:, dy:dy + H, dx:dx + W]
    return out


def contour_complexity_tiles(binmask: np.ndarray, tile: int, detail: dict | None = None):
    """morphology.py:709-739 with contour_components=True."""
    eroded = erode3(binmask)
    boundary = binmask & ~eroded
    area_i = tile_count(binmask, tile)
    perim_i = tile_count(boundary, tile)
    e4 = euler_x4_tiles(binmask, tile)
    if detail is not None:
        detail.update(area=area_i, perim=perim_i, euler_x4=e4)
    area = area_i.astype(f32)
    perim = perim_i.astype(f32)
    ic = ((perim * perim).astype(f32) / ((FOUR_PI * area).astype(f32) + f32(1e-6))).astype(f32)
    K = np.maximum(rint32(e4.astype(f32) / f32(4.0)), f32(1.0))
    ic = (ic / K).astype(f32)
    phi5 = (f32(1.0) - (f32(1.0) / np.maximum(ic, f32(1.0))).astype(f32)).astype(f32)
    return np.where(area_i > 0, phi5, f32(0.0)).astype(f32)


def phi_tiles_from_sum(sum_plane: np.ndarray, C: int, grid_size: int = 8, detail: dict | None = None):
    """morphology.py:826-873 given the channel-sum plane (B,H,W)."""
    B, H, W = sum_plane.shape
    tile = tile_size(H, grid_size)
    ht, wt = H // tile, W // tile
    Hc, Wc = ht * tile, wt * tile
    gray = (sum_plane[:, :Hc, :Wc] / f32(C)).astype(f32)
    gray = normalize01(gray)
    gx, gy = sobel(gray)
    d_c = {} if detail is not None else None
    edge = canny_edges(gray, detail=d_c)
    binmask = adaptive_binarize(gray, detail=d_c)
    phi1 = (fractal_dimension_tiles(edge, tile, d_c) / f32(2.0)).astype(f32)
    phi2 = lbp_entropy_tiles(gray, tile, d_c)
    phi3 = gradient_variance_tiles(gx, gy, tile)
    edge_cnt = tile_count(edge, tile)
    phi4 = (edge_cnt.astype(f32) / f32(tile * tile)).astype(f32)
    phi5 = contour_complexity_tiles(binmask, tile, d_c)
    phi = np.stack(
        [phi1, phi2, phi3, phi4, phi5,
         (phi1 * phi2).astype(f32), (phi3 * phi3).astype(f32),
         sqrt32((phi4 * phi5).astype(f32) + f32(1e-12))], axis=-1).astype(f32)
    if detail is not None:
        detail.update(d_c)
        detail.update(gray=gray, gx=gx, gy=gy, edge=edge, binmask=binmask, tile=tile,
                      edge_count=edge_cnt)
    return phi


def phi_tiles(x: np.ndarray, grid_size: int = 8, detail: dict | None = None):
    s, _ = channel_sums(x)
    return phi_tiles_from_sum(s, x.shape[1], grid_size, detail)


# ----------------------------------------------------------------------------
# small dense layers
# ----------------------------------------------------------------------------

def linear(x: np.ndarray, W: np.ndarray, b: np.ndarray) -> np.ndarray:
    """nn.Linear: sequential FMA over k from 0, bias last.  x (N,K), W (M,K)."""
    N, K = x.shape
    acc = np.zeros((N, W.shape[0]), dtype=f32)
    for k in range(K):
        acc = fma32(x[:, k:k + 1], W[:, k][None, :], acc)
    return (acc + b[None, :]).astype(f32)


def tree_sum32(v: np.ndarray) -> np.ndarray:
    """Sum of 32 values per row in xor-butterfly order (step 16, 8, 4, 2, 1) -- the order a warp
    shuffle reduction produces; all lanes end with identical bits.  v: (N,32) -> (N,)."""
    idx = np.arange(32)
    v = v.astype(f32)
    for o in (16, 8, 4, 2, 1):
        v = (v + v[:, idx ^ o]).astype(f32)
    return v[:, 0]


def layer_norm(x: np.ndarray, g: np.ndarray, b: np.ndarray, eps: float = 1e-5):
    """nn.LayerNorm over the last axis (D = 32 or 64).  Statistics are reduced with a 32-lane
    butterfly (elements k and k+32 pre-added for D = 64); torch's own CPU order (vectorised
    Welford) is not reproducible outside torch, the difference stays at the 1e-7 level."""
    N, D = x.shape
    assert D in (32, 64), D

    def fold(a):
        return a.astype(f32) if D == 32 else (a[:, :32] + a[:, 32:]).astype(f32)

    mean = (tree_sum32(fold(x)) / f32(D)).astype(f32)
    d = (x - mean[:, None]).astype(f32)
    var = (tree_sum32(fold((d * d).astype(f32))) / f32(D)).astype(f32)
    rstd = (f32(1.0) / sqrt32(var + f32(eps))).astype(f32)
    return (((d * rstd[:, None]).astype(f32) * g[None, :]).astype(f32) + b[None, :]).astype(f32)


def sigmoid32(x):
    return (f32(1.0) / (f32(1.0) + exp32(-x))).astype(f32)


def complexity_mlp(phi: np.ndarray, w: dict) -> np.ndarray:
    """morphology.py:81-90.  phi (N,8) -> (N,).  w keys: state_dict names
    `complexity_mlp.{0,3,6}.{weight,bias}`, `complexity_mlp.{1,4}.{weight,bias}`."""
    h = linear(phi, w["complexity_mlp.0.weight"], w["complexity_mlp.0.bias"])
    h = np.maximum(layer_norm(h, w["complexity_mlp.1.weight"], w["complexity_mlp.1.bias"]), f32(0))
    h = linear(h, w["complexity_mlp.3.weight"], w["complexity_mlp.3.bias"])
    h = np.maximum(layer_norm(h, w["complexity_mlp.4.weight"], w["complexity_mlp.4.bias"]), f32(0))
    h = linear(h, w["complexity_mlp.6.weight"], w["complexity_mlp.6.bias"])
    return sigmoid32(h[:, 0])


def bilateral_filter(cmap: np.ndarray) -> np.ndarray:
    """morphology.py:309-354 (5x5, sigma_s=2, sigma_r=0.1, replicate pad)."""
    B, H, W = cmap.shape
    cp = np.pad(cmap, ((0, 0), (2, 2), (2, 2)), mode="edge")
    num = np.zeros((B, H, W), dtype=f32)
    den = np.zeros((B, H, W), dtype=f32)
    for ky in range(5):
        for kx in range(5):
            p = cp[:, ky:ky + H, kx:kx + W]
            d = (p - cmap).astype(f32)
            rw = exp32((-(d * d).astype(f32)) / BILATERAL_RANGE_DEN)
            wgt = (BILATERAL_SPATIAL[ky, kx] * rw).astype(f32)
            num = (num + (wgt * p).astype(f32)).astype(f32)
            den = (den + wgt).astype(f32)
    return (num / (den + f32(1e-8))).astype(f32)


def analyzer_forward(x: np.ndarray, w: dict, grid_size: int = 8, detail: dict | None = None):
    """MorphologicalComplexityAnalyzer.forward (morphology.py:939-973)."""
    phi = phi_tiles(x, grid_size, detail)
    return complexity_from_phi(phi, w, detail)


def complexity_from_phi(phi: np.ndarray, w: dict, detail: dict | None = None):
    B, ht, wt, _ = phi.shape
    raw = complexity_mlp(phi.reshape(-1, 8), w).reshape(B, ht, wt)
    c = np.clip(bilateral_filter(raw), f32(0.0), f32(1.0)).astype(f32)
    if detail is not None:
        detail.update(phi=phi, complexity_raw=raw)
    return c


def score_image(x: np.ndarray, feature_weights: np.ndarray, grid_size: int = 8):
    """morphology.py:923-937."""
    phi = phi_tiles(x, grid_size)
    a = np.abs(feature_weights.astype(f32))
    tot = f32(0)
    for v in a:
        tot = f32(tot + v)
    a = (a / f32(max(tot, f32(1e-8)))).astype(f32)
    c = np.zeros(phi.shape[:3], dtype=f32)
    for k in range(5):
        c = (c + (phi[..., k] * a[k]).astype(f32)).astype(f32)
    B = c.shape[0]
    flat = c.reshape(B, -1)
    s = np.zeros(B, dtype=f32)
    for k in range(flat.shape[1]):
        s = s + flat[:, k]
    return np.clip((s / f32(flat.shape[1])).astype(f32), f32(0), f32(1))


# ----------------------------------------------------------------------------
# bit mappers
# ----------------------------------------------------------------------------

def _finish_bits(b: np.ndarray, temperature, lo: float, hi: float, continuous: bool):
    """temperature multiply, STE clamp, STE round (bit_allocation.py:72-80, 264-280)."""
    if temperature is not None:
        b = (b * f32(max(float(temperature), 0.1))).astype(f32)
    clamped = np.clip(b, f32(lo), f32(hi))
    b = (b + (clamped - b).astype(f32)).astype(f32)
    if not continuous:
        b = (b + (rint32(b) - b).astype(f32)).astype(f32)
    return b


def batch_norm_eval(x, g, b, rm, rv, eps: float = 1e-5):
    invstd = (1.0 / np.sqrt(rv.astype(f64) + eps)).astype(f32)
    alpha = (invstd * g).astype(f32)
    beta = (b - (rm * alpha).astype(f32)).astype(f32)
    return ((x * alpha[None, :]).astype(f32) + beta[None, :]).astype(f32)


def mlp_bit_mapper(c: np.ndarray, w: dict, temperature=1.0, continuous: bool = False,
                   min_bits: float = 2.0, max_bits: float = 8.0, detail: dict | None = None):
    """ComplexityToBitMappingNetwork.forward in eval mode (bit_allocation.py:218-280).
    w keys: `mapping_network.{0,3,6,9}.{weight,bias}`,
    `mapping_network.{1,4,7}.{weight,bias,running_mean,running_var}`."""
    B, H, W = c.shape
    cc = np.clip(c, f32(0), f32(1)).astype(f32).reshape(-1)
    z = np.stack([cc, (cc * cc).astype(f32), log1p32(cc)], axis=1)
    h = z
    for li, bi in ((0, 1), (3, 4), (6, 7)):
        h = linear(h, w[f"mapping_network.{li}.weight"], w[f"mapping_network.{li}.bias"])
        h = batch_norm_eval(h, w[f"mapping_network.{bi}.weight"], w[f"mapping_network.{bi}.bias"],
                            w[f"mapping_network.{bi}.running_mean"], w[f"mapping_network.{bi}.running_var"])
        h = np.maximum(h, f32(0))
    logit = linear(h, w["mapping_network.9.weight"], w["mapping_network.9.bias"])[:, 0]
    s = sigmoid32(logit)
    b = (f32(min_bits) + (f32(max_bits - min_bits) * s).astype(f32)).astype(f32).reshape(B, H, W)
    if detail is not None:
        pre = b if temperature is None else (b * f32(max(float(temperature), 0.1))).astype(f32)
        detail.update(bits_pre_round=np.clip(pre, f32(min_bits), f32(max_bits)), mapper_logit=logit)
    return _finish_bits(b, temperature, min_bits, max_bits, continuous)


def quantile_lerp(flat: np.ndarray, q: float) -> np.ndarray:
    """torch.quantile(linear) over the last axis: fp32 rank, torch.lerp formula."""
    n = flat.shape[1]
    srt = np.sort(flat, axis=1)
    rank = f32(f32(q) * f32(n - 1))
    lo = int(math.floor(rank))
    hi = int(math.ceil(rank))
    wgt = f32(rank - f32(lo))
    a, b = srt[:, lo], srt[:, hi]
    diff = (b - a).astype(f32)
    if wgt < f32(0.5):
        return (a + (wgt * diff).astype(f32)).astype(f32)
    return (b - (diff * f32(f32(1) - wgt)).astype(f32)).astype(f32)


def linear_bit_mapper(c: np.ndarray, temperature=1.0, continuous: bool = False,
                      min_bits: float = 2.0, max_bits: float = 8.0, eps_spread: float = 1e-3,
                      detail: dict | None = None):
    """LinearBitMapper.forward (bit_allocation.py:42-80)."""
    B = c.shape[0]
    flat = c.reshape(B, -1).astype(f32)
    lo = quantile_lerp(flat, 0.02)[:, None, None]
    hi = quantile_lerp(flat, 0.98)[:, None, None]
    spread = (hi - lo).astype(f32)
    rel = np.clip(((c - lo).astype(f32) / (spread + f32(1e-8))).astype(f32), f32(0), f32(1))
    cn = np.where(spread > f32(eps_spread), rel, np.clip(c, f32(0), f32(1))).astype(f32)
    b = (f32(min_bits) + (f32(max_bits - min_bits) * cn).astype(f32)).astype(f32)
    if detail is not None:
        pre = b if temperature is None else (b * f32(max(float(temperature), 0.1))).astype(f32)
        detail.update(bits_pre_round=np.clip(pre, f32(min_bits), f32(max_bits)), q_lo=lo, q_hi=hi)
    return _finish_bits(b, temperature, min_bits, max_bits, continuous)


def percentile_normalize(c: np.ndarray) -> np.ndarray:
    """Optional hook-level normalisation (models/mcaq_yolo.py:427-432)."""
    B = c.shape[0]
    flat = c.reshape(B, -1).astype(f32)
    lo = quantile_lerp(flat, 0.02)[:, None, None]
    hi = quantile_lerp(flat, 0.98)[:, None, None]
    return np.clip(((c - lo).astype(f32) / ((hi - lo).astype(f32) + f32(1e-8))).astype(f32), f32(0), f32(1))


# ----------------------------------------------------------------------------
# learned soft mask  (quantization.py:168-239)
# ----------------------------------------------------------------------------

def soft_mask_tiles(bit_map: np.ndarray, abs_sum: np.ndarray, C: int, w: dict,
                    detail: dict | None = None) -> np.ndarray:
    """Tile-level mask value m_T (B,Ht,Wt) before upsampling.
    abs_sum: sum_c |x| (B,H,W).  w keys: `soft_mask.net.{0,2}.{weight,bias}`."""
    B, H, W = abs_sum.shape
    Ht, Wt = bit_map.shape[-2:]
    amean = (abs_sum / f32(C)).astype(f32)
    ys, ye = adaptive_windows(H, Ht)
    xs, xe = adaptive_windows(W, Wt)
    act = np.zeros((B, Ht, Wt), dtype=f32)
    for i in range(Ht):
        for j in range(Wt):
            win = amean[:, ys[i]:ye[i], xs[j]:xe[j]]
            cols = np.zeros((B, win.shape[2]), dtype=f32)
            for y in range(win.shape[1]):                     # column sums top-to-bottom ...
                cols = cols + win[:, y, :]
            tot = np.zeros(B, dtype=f32)
            for x in range(win.shape[2]):                     # ... then left-to-right
                tot = tot + cols[:, x]
            act[:, i, j] = tot / f32(win.shape[1] * win.shape[2])
    act = (act / (act.max(axis=(1, 2), keepdims=True) + f32(1e-8))).astype(f32)
    bits_norm = np.clip(((bit_map.astype(f32) - f32(2.0)) / f32(6.0)).astype(f32), f32(0), f32(1))
    feats = np.stack([bits_norm, act], axis=1)                       # (B,2,Ht,Wt)
    W0, b0 = w["soft_mask.net.0.weight"], w["soft_mask.net.0.bias"]   # (8,2,3,3)
    W2, b2 = w["soft_mask.net.2.weight"], w["soft_mask.net.2.bias"]   # (2,8,1,1)
    fp = np.pad(feats, ((0, 0), (0, 0), (1, 1), (1, 1)))
    hid = np.zeros((B, W0.shape[0], Ht, Wt), dtype=f32)
    for o in range(W0.shape[0]):
        acc = np.zeros((B, Ht, Wt), dtype=f32)
        for ic in range(2):
            for ky in range(3):
                for kx in range(3):
                    acc = fma32(fp[:, ic, ky:ky + Ht, kx:kx + Wt], W0[o, ic, ky, kx], acc)
        hid[:, o] = np.maximum((acc + b0[o]).astype(f32), f32(0))
    logits = np.zeros((B, 2, Ht, Wt), dtype=f32)
    for o in range(2):
        acc = np.zeros((B, Ht, Wt), dtype=f32)
        for ic in range(W0.shape[0]):
            acc = fma32(hid[:, ic], W2[o, ic, 0, 0], acc)
        logits[:, o] = (acc + b2[o]).astype(f32)
    mx = np.maximum(logits[:, 0], logits[:, 1])
    e0 = exp32((logits[:, 0] - mx).astype(f32))
    e1 = exp32((logits[:, 1] - mx).astype(f32))
    mt = (e0 / (e0 + e1).astype(f32)).astype(f32)
    if detail is not None:
        detail.update(act=act, mask_tiles=mt)
    return mt


def soft_mask(bit_map: np.ndarray, abs_sum: np.ndarray, C: int, w: dict,
              detail: dict | None = None) -> np.ndarray:
    """LearnedSoftMask.forward -> m (B,H,W).  `soft_mask.smooth_kernel` (1,1,5,5)."""
    B, H, W = abs_sum.shape
    Ht, Wt = bit_map.shape[-2:]
    mt = soft_mask_tiles(bit_map, abs_sum, C, w, detail)
    iy = nearest_index(H, Ht)
    ix = nearest_index(W, Wt)
    up = mt[:, iy][:, :, ix]
    k = w["soft_mask.smooth_kernel"].reshape(5, 5).astype(f32)
    return stencil(up, k, "edge")


# ----------------------------------------------------------------------------
# quantizer  (K3)
# ----------------------------------------------------------------------------

def qparams(mn: np.ndarray, mx: np.ndarray, bits: int):
    """QuantizationParameters.compute_scale_zeropoint (quantization.py:41-66)."""
    qmin = -(2 ** (bits - 1))
    qmax = 2 ** (bits - 1) - 1
    rng = np.maximum((mx - mn).astype(f32), f32(1e-8))
    scale = (rng / f32(qmax - qmin)).astype(f32)
    zp = (f32(qmin) - (mn / scale).astype(f32)).astype(f32)
    zp = np.clip(zp, f32(qmin), f32(qmax)).astype(f32)
    return scale, zp, qmin, qmax


def quant_dequant(x: np.ndarray, mn, mx, bits: int):
    """quantize_tensor (quantization.py:597-600) for per-channel ranges.
    Returns (dequantised fp32, integer codes int16)."""
    scale, zp, qmin, qmax = qparams(mn, mx, bits)
    s = scale[None, :, None, None]
    z = zp[None, :, None, None]
    q = rint32(((x / s).astype(f32) + z).astype(f32))
    q = np.clip(q, f32(qmin), f32(qmax))
    return ((q - z).astype(f32) * s).astype(f32), q.astype(np.int16)


def tile_lookup(H: int, W: int, Ht: int, Wt: int):
    return nearest_index(H, Ht), nearest_index(W, Wt)


def quantize_eval(x: np.ndarray, bit_map: np.ndarray, mn, mx, m: np.ndarray | None):
    """Inference compose (quantization.py:729-744).  bit_map holds integers.
    Returns (y fp32, codes int16)."""
    x = np.asarray(x, dtype=f32)
    B, C, H, W = x.shape
    Ht, Wt = bit_map.shape[-2:]
    iy, ix = tile_lookup(H, W, Ht, Wt)
    bpix = bit_map[:, iy][:, :, ix]                               # (B,H,W)
    y = np.zeros_like(x)
    codes = np.zeros(x.shape, dtype=np.int16)
    for bv in np.unique(bit_map):
        bits = int(round(float(bv)))
        dq, q = quant_dequant(x, mn, mx, bits)
        sel = (bpix == bv)[:, None, :, :]
        y = np.where(sel, dq, y)
        codes = np.where(sel, q, codes)
    if m is not None:
        y = (y * m[:, None, :, :]).astype(f32)
    return y.astype(f32), codes


def quantize_train_fwd(x: np.ndarray, bit_map: np.ndarray, mn, mx, m: np.ndarray | None):
    """Fractional-bit training compose (quantization.py:699-727, 742-744).
    Returns (y, pre, q_lo, q_hi, frac_pix)."""
    x = np.asarray(x, dtype=f32)
    B, C, H, W = x.shape
    Ht, Wt = bit_map.shape[-2:]
    iy, ix = tile_lookup(H, W, Ht, Wt)
    bfl = np.floor(bit_map.astype(f32))
    frac = (bit_map.astype(f32) - bfl).astype(f32)
    fpix = frac[:, iy][:, :, ix][:, None]
    bpix = bfl[:, iy][:, :, ix][:, None]
    qlo = np.zeros_like(x)
    qhi = np.zeros_like(x)
    for bf in np.unique(bfl):
        b0 = int(bf)
        lo, _ = quant_dequant(x, mn, mx, b0)
        hi = quant_dequant(x, mn, mx, b0 + 1)[0] if b0 + 1 <= 8 else lo
        sel = bpix == bf
        qlo = np.where(sel, lo, qlo)
        qhi = np.where(sel, hi, qhi)
    pre = (((f32(1.0) - fpix).astype(f32) * qlo).astype(f32) + (fpix * qhi).astype(f32)).astype(f32)
    y = pre if m is None else (pre * m[:, None]).astype(f32)
    return y.astype(f32), pre, qlo, qhi, fpix[:, 0]


def quantize_train_bwd(g: np.ndarray, x, bit_map, mn, mx, m):
    """Analytic backward of the training compose:
    dx = g*m*(1-f) + g*m*f ; dbit[tile] = sum g*m*(q_hi-q_lo) ; dm = sum_c g*pre.
    Accumulations in fp64 (order-free reference for a tolerance check)."""
    y, pre, qlo, qhi, fpix = quantize_train_fwd(x, bit_map, mn, mx, m)
    B, C, H, W = x.shape
    Ht, Wt = bit_map.shape[-2:]
    mm = np.ones((B, 1, H, W), dtype=f32) if m is None else m[:, None]
    gm = (g * mm).astype(f32)
    f = fpix[:, None]
    dx = ((gm * (f32(1.0) - f).astype(f32)).astype(f32) + (gm * f).astype(f32)).astype(f32)
    iy, ix = tile_lookup(H, W, Ht, Wt)
    contrib = (gm.astype(f64) * (qhi - qlo).astype(f64)).sum(axis=1)       # (B,H,W)
    dbit = np.zeros((B, Ht, Wt), dtype=f64)
    np.add.at(dbit, (np.arange(B)[:, None, None], iy[None, :, None], ix[None, None, :]), contrib)
    dm = (g.astype(f64) * pre.astype(f64)).sum(axis=1)
    return dx, dbit, dm


# ----------------------------------------------------------------------------
# hook composition  (models/mcaq_yolo.py:409-455, inference)
# ----------------------------------------------------------------------------

def hook_forward(x: np.ndarray, analyzer_w: dict, mapper_w: dict | None, quant_w: dict | None,
                 grid_size: int = 8, temperature: float = 1.0, frozen_minmax=None,
                 detail: dict | None = None):
    """One scale of the inference hook: analyzer -> mapper -> quantizer.
    mapper_w None selects LinearBitMapper; quant_w None disables the soft mask.
    Returns dict(complexity, bit_map, m, y, codes, min, max)."""
    x = np.asarray(x, dtype=f32)
    B, C, H, W = x.shape
    s, a = channel_sums(x)
    phi = phi_tiles_from_sum(s, C, grid_size, detail)
    c = complexity_from_phi(phi, analyzer_w, detail)
    if mapper_w is None:
        bm = linear_bit_mapper(c, temperature, False, detail=detail)
    else:
        bm = mlp_bit_mapper(c, mapper_w, temperature, False, detail=detail)
    mn, mx = channel_minmax(x) if frozen_minmax is None else frozen_minmax
    m = soft_mask(bm, a, C, quant_w, detail) if quant_w is not None else None
    y, codes = quantize_eval(x, bm, mn, mx, m)
    return dict(phi=phi, complexity=c, bit_map=bm, m=m, y=y, codes=codes, min=mn, max=mx,
                sum=s, abs_sum=a)
