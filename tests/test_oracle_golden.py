"""Pins the numpy oracle (oracle/mcaq_oracle.py) to outputs of the REAL reference stored in
tests/golden/ (tools/make_golden.py ran /root/reference on the same integer-built inputs).

Bar: bit-exact for every discrete quantity (edge / binary planes, bit maps, integer codes);
rtol 1e-4 for floating-point ones (north_star), with the tolerance written in each test.
"""
import numpy as np
import pytest

import mcaq_oracle as o
from golden_util import CASE_NAMES, Case, bit_ambiguous, sha, weights
from inputs import fractional_bit_map, integer_bit_map

RTOL = 1e-4          # north_star: complexity metrics / de-quantised maps within rtol 1e-4 (fp32)
ATOL = 2e-6          # a few fp32 ulps at |v| <= 8 for quantities that pass through exp/log


@pytest.fixture(scope="module")
def W():
    return weights()


def test_constants_match_torch(W):
    """Stencil taps and fractal tables the oracle (and the CUDA side) use are exactly torch's."""
    assert np.array_equal(W["const"]["canny_blur"], o.CANNY_BLUR)
    assert np.array_equal(W["const"]["adapt_blur"], o.ADAPT_BLUR)
    assert np.array_equal(W["const"]["bilateral_spatial"], o.BILATERAL_SPATIAL)
    assert np.array_equal(o.gaussian_kernel2d(5, 1.0), o.CANNY_BLUR)
    assert np.array_equal(o.gaussian_kernel2d(5, 5 / 3.0),
                          W["quantizer"]["soft_mask.smooth_kernel"].reshape(5, 5))
    for t in (4, 8, 16, 32):
        _, x, w = o.fractal_tables(t)
        assert np.array_equal(x, W["const"][f"frac_logs_{t}"])
        assert np.array_equal(w, W["const"][f"frac_w_{t}"])


@pytest.fixture(scope="module", params=CASE_NAMES)
def run(request, W):
    c = Case(request.param)
    d = {}
    r = o.hook_forward(c.x(), W["analyzer"], W["mapper"], W["quantizer"], c.grid, 1.0, detail=d)
    return c, r, d


def test_tile_geometry(run):
    c, r, d = run
    assert d["tile"] == c.tile and r["phi"].shape == (c.B, c.ht, c.wt, 8)


def test_gray_plane_bit_exact(run):
    """Cascade channel sum + per-image min-max normalise (morphology.py:837-838)."""
    c, r, d = run
    assert np.array_equal(d["gray"], c["gray"])


def test_edge_and_binary_planes_bit_exact(run):
    """Canny (blur, Otsu, Sobel, NMS, hysteresis) and adaptive threshold planes."""
    c, r, d = run
    assert int((d["edge"] != c.plane_bits("edge")).sum()) == 0
    assert int((d["binmask"] != c.plane_bits("binmask")).sum()) == 0


def test_phi_and_complexity(run):
    c, r, d = run
    np.testing.assert_allclose(r["phi"], c["phi"], rtol=RTOL, atol=ATOL)
    # phi1 / phi4 / phi5 are functions of integer counts only -> exact
    for k in (0, 3, 4):
        assert np.array_equal(r["phi"][..., k], c["phi"][..., k]), k
    np.testing.assert_allclose(d["complexity_raw"], c["complexity_raw"], rtol=RTOL, atol=ATOL)
    np.testing.assert_allclose(r["complexity"], c["complexity"], rtol=RTOL, atol=ATOL)


def test_bit_maps_bit_exact(run, W):
    c, r, d = run
    amb = bit_ambiguous(d["bits_pre_round"])
    neq = r["bit_map"] != c["bit_map_mlp"]
    assert int((neq & ~amb).sum()) == 0, f"{int(neq.sum())} tiles differ, {int(amb.sum())} ambiguous"
    assert int(neq.sum()) == 0          # holds for every committed case today
    d2 = {}
    lin = o.linear_bit_mapper(c["complexity"], 1.0, False, detail=d2)
    assert np.array_equal(lin, c["bit_map_linear"])
    cont = o.mlp_bit_mapper(c["complexity"], W["mapper"], 1.3, True)
    np.testing.assert_allclose(cont, c["bit_map_mlp_cont_T13"], rtol=RTOL, atol=ATOL)


def test_soft_mask(run):
    c, r, d = run
    np.testing.assert_allclose(r["m"], c["soft_mask"], rtol=RTOL, atol=ATOL)


def test_ranges_and_codes_bit_exact(run, W):
    c, r, d = run
    assert np.array_equal(r["min"], c["ch_min"]) and np.array_equal(r["max"], c["ch_max"])
    assert sha(r["codes"]) == str(c["codes_sha"])
    assert np.array_equal(r["codes"][:, ::3, ::5, ::7], c["codes_sub"])
    # de-quantised map: 1e-4 relative (differs from the reference only through m's last ulps)
    np.testing.assert_allclose(r["y"][:, ::3, ::5, ::7], c["y_sub"], rtol=RTOL, atol=ATOL)
    # all widths 2..8: random integer map, the reference's own mask -> bit-identical y
    x = c.x()
    bm = integer_bit_map(c.B, c.ht, c.wt, c.seed)
    assert np.array_equal(bm, c["bit_map_rand"])
    _, a = o.channel_sums(x)
    m = o.soft_mask(bm, a, c.C, W["quantizer"])
    y, codes = o.quantize_eval(x, bm, r["min"], r["max"], m)
    assert sha(codes) == str(c["codes_rand_sha"])
    np.testing.assert_allclose(y[:, ::3, ::5, ::7], c["y_rand_sub"], rtol=RTOL, atol=ATOL)


def test_score_image(run, W):
    c, r, d = run
    s = o.score_image(c.x(), W["analyzer"]["feature_weights"], c.grid)
    np.testing.assert_allclose(s, c["score"], rtol=RTOL, atol=ATOL)


def test_training_path(run, W):
    """Fractional-bit forward + STE backward (quantization.py:699-727, 101-118)."""
    c, r, d = run
    x, g = c.x(), c.grad()
    bf = fractional_bit_map(c.B, c.ht, c.wt, c.seed)
    assert np.array_equal(bf, c["bit_map_frac"])
    mn, mx = o.ema_update(None, None, *o.channel_minmax(x))
    assert np.array_equal(mn, c["train_run_min"]) and np.array_equal(mx, c["train_run_max"])
    # with the reference's own mask the training forward and dx are bit-identical
    m_ref = c["train_soft_mask"]
    y, pre, qlo, qhi, fp = o.quantize_train_fwd(x, bf, mn, mx, m_ref)
    assert sha(y) == str(c["train_y_sha"])
    dx, dbit, dm = o.quantize_train_bwd(g, x, bf, mn, mx, m_ref)
    assert sha(dx) == str(c["train_dx_sha"])
    # oracle's own mask: within tolerance
    _, a = o.channel_sums(x)
    m = o.soft_mask(bf, a, c.C, W["quantizer"])
    np.testing.assert_allclose(m, m_ref, rtol=RTOL, atol=ATOL)
    # mask-off: d(bit_map) is exactly the fractional term (fp32 autograd sum vs fp64 here)
    y2, *_ = o.quantize_train_fwd(x, bf, mn, mx, None)
    assert sha(y2) == str(c["train_nomask_y_sha"])
    _, dbit2, _ = o.quantize_train_bwd(g, x, bf, mn, mx, None)
    ref = c["train_nomask_dbit"]
    np.testing.assert_allclose(dbit2, ref, rtol=2e-3, atol=2e-3 * np.abs(ref).max())
    # second EMA step (momentum 0.99)
    x2 = (x * np.float32(1.5) + np.float32(0.25)).astype(np.float32)
    mn2, mx2 = o.ema_update(mn, mx, *o.channel_minmax(x2))
    np.testing.assert_allclose(mn2, c["ema2_min"], rtol=1e-6)
    np.testing.assert_allclose(mx2, c["ema2_max"], rtol=1e-6)


# ---- the reference's own known-answer tests for this path (tests/test_smoke.py) ------------

def test_euler_component_count_known_answer():
    """tests/test_smoke.py:214-223."""
    m = np.zeros((1, 16, 16), dtype=bool)
    m[0, 2:6, 2:6] = True
    assert o.euler_x4_tiles(m, 16)[0, 0, 0] == 4
    m[0, 10:14, 10:14] = True
    assert o.euler_x4_tiles(m, 16)[0, 0, 0] == 8


def test_linear_mapper_known_answers():
    """tests/test_smoke.py:188-211."""
    c = (np.linspace(0, 1, 16, dtype=np.float32).reshape(1, 4, 4) * np.float32(0.05)
         + np.float32(0.4)).astype(np.float32)
    b = o.linear_bit_mapper(c, 1.0)
    assert b.min() == 2.0 and b.max() == 8.0 and np.unique(b).size >= 5
    assert np.all(o.linear_bit_mapper(np.full((1, 8, 8), 0.5, np.float32)) == 5)
    assert o.linear_bit_mapper(np.full((1, 8, 8), 0.0, np.float32)).max() == 2
    assert o.linear_bit_mapper(np.full((1, 8, 8), 1.0, np.float32)).min() == 8


def test_mapper_temperature_saturates(W):
    """tests/test_smoke.py:74-84: integer bits in [2,8]; T=10 -> all 8."""
    rng = np.random.Generator(np.random.PCG64(7))
    c = rng.random((2, 8, 8), dtype=np.float32)
    b = o.mlp_bit_mapper(c, W["mapper"], 1.0)
    assert b.min() >= 2 and b.max() <= 8 and np.array_equal(b, np.rint(b))
    assert np.all(o.mlp_bit_mapper(c, W["mapper"], 10.0) == 8.0)


def test_phi_range_all_sizes():
    """tests/test_smoke.py:33-47 (shapes, pow-2 tile >= 4, phi in [0,1])."""
    rng = np.random.Generator(np.random.PCG64(3))
    for H in (160, 80, 40, 20):
        x = rng.random((2, 3, H, H), dtype=np.float32)
        phi = o.phi_tiles(x, 8)
        t = o.tile_size(H, 8)
        assert phi.shape == (2, H // t, H // t, 8) and t >= 4 and (t & (t - 1)) == 0
        assert phi.min() >= 0.0 and phi.max() <= 1.0 + 1e-5


def test_fma32_is_single_rounding():
    rng = np.random.Generator(np.random.PCG64(5))
    a = rng.standard_normal(200000).astype(np.float32)
    b = rng.standard_normal(200000).astype(np.float32)
    c = rng.standard_normal(200000).astype(np.float32)
    from fractions import Fraction
    r = o.fma32(a, b, c)
    for i in range(0, 200000, 997):
        exact = Fraction(float(a[i])) * Fraction(float(b[i])) + Fraction(float(c[i]))
        lo = np.nextafter(r[i], np.float32(-np.inf))
        hi = np.nextafter(r[i], np.float32(np.inf))
        assert abs(Fraction(float(r[i])) - exact) <= min(abs(Fraction(float(lo)) - exact),
                                                         abs(Fraction(float(hi)) - exact))
