"""GPU parity: every kernel, called through the C ABI (ctypes), against the numpy oracle on the
seeded golden-case inputs, and against the committed reference outputs (tests/golden).

Bar (north_star): bit maps and integer codes bit-exact; complexity metrics and de-quantised
maps within rtol 1e-4 (fp32) / 1 LSB of the assigned width (bf16).
"""
import numpy as np
import pytest
import torch

import mcaq_oracle as o
from golden_util import CASE_NAMES, Case, bit_ambiguous, sha, weights

pytestmark = pytest.mark.gpu

RTOL = 1e-4
ATOL = 2e-6


@pytest.fixture(scope="module")
def ops():
    from mcaq_yolo_b200 import ops as _ops
    return _ops


@pytest.fixture(scope="module")
def K():
    from mcaq_yolo_b200 import constants
    return constants


@pytest.fixture(scope="module")
def W():
    return weights()


def dev(a, dtype=None):
    t = torch.from_numpy(np.ascontiguousarray(a)).cuda()
    return t if dtype is None else t.to(dtype)


def unpack_bits(words: torch.Tensor, Wc: int) -> np.ndarray:
    w = words.cpu().numpy().astype(np.uint32)
    bits = ((w[..., None] >> np.arange(32, dtype=np.uint32)) & 1).astype(bool)
    return bits.reshape(w.shape[0], w.shape[1], -1)[:, :, :Wc]


def packed_params(W, K):
    """Device parameter blocks built from the golden fixtures' state_dicts, in the layout
    include/mcaq_b200.h documents (weights transposed to [k][unit], eval BN folded, zero pads)."""
    a = W["analyzer"]
    z3 = np.zeros(3, np.float32)
    cm = np.concatenate([a["complexity_mlp.0.weight"].T.ravel(), a["complexity_mlp.0.bias"],
                         a["complexity_mlp.1.weight"], a["complexity_mlp.1.bias"],
                         a["complexity_mlp.3.weight"].T.ravel(), a["complexity_mlp.3.bias"],
                         a["complexity_mlp.4.weight"], a["complexity_mlp.4.bias"],
                         a["complexity_mlp.6.weight"].ravel(), a["complexity_mlp.6.bias"].ravel(), z3])
    m = W["mapper"]
    parts = []
    for li, bi in ((0, 1), (3, 4), (6, 7)):
        invstd = (1.0 / np.sqrt(m[f"mapping_network.{bi}.running_var"].astype(np.float64) + 1e-5)).astype(np.float32)
        alpha = (invstd * m[f"mapping_network.{bi}.weight"]).astype(np.float32)
        beta = (m[f"mapping_network.{bi}.bias"] - (m[f"mapping_network.{bi}.running_mean"] * alpha).astype(np.float32)).astype(np.float32)
        parts += [m[f"mapping_network.{li}.weight"].T.ravel(), m[f"mapping_network.{li}.bias"].ravel(), alpha, beta]
    parts += [m["mapping_network.9.weight"].ravel(), m["mapping_network.9.bias"].ravel(), z3]
    mp = np.concatenate(parts)
    q = W["quantizer"]
    sm = np.concatenate([q["soft_mask.net.0.weight"].ravel(), q["soft_mask.net.0.bias"].ravel(),
                         q["soft_mask.net.2.weight"].ravel(), q["soft_mask.net.2.bias"].ravel(),
                         q["soft_mask.smooth_kernel"].ravel(), z3[:1]])
    assert cm.size == K.CMLP_FLOATS and mp.size == K.MAPPER_FLOATS and sm.size == K.SOFTMASK_FLOATS
    return dev(cm.astype(np.float32)), dev(mp.astype(np.float32)), dev(sm.astype(np.float32))


@pytest.fixture(scope="module", params=CASE_NAMES)
def case(request, W):
    c = Case(request.param)
    d = {}
    x = c.x()
    r = o.hook_forward(x, W["analyzer"], W["mapper"], W["quantizer"], c.grid, 1.0, detail=d)
    return c, x, r, d


# ------------------------------------------------------------------------------------------ K1
def test_reduce_planes_bit_exact(case, ops):
    c, x, r, d = case
    s, a, keys = ops.reduce_planes(dev(x))
    assert np.array_equal(s.cpu().numpy(), r["sum"]), "channel sum (cascade order) differs"
    assert np.array_equal(a.cpu().numpy(), r["abs_sum"])
    packed = ops.ranges_decode(keys).cpu().numpy()
    assert np.array_equal(packed[:c.C], r["min"]) and np.array_equal(-packed[c.C:], r["max"])


def test_reduce_planes_bf16(case, ops):
    c, x, r, d = case
    xb = dev(x, torch.bfloat16)
    xu = xb.float().cpu().numpy()                       # oracle sees the exact upcast values
    s, a, keys = ops.reduce_planes(xb)
    so, ao = o.channel_sums(xu)
    assert np.array_equal(s.cpu().numpy(), so) and np.array_equal(a.cpu().numpy(), ao)
    mn, mx = o.channel_minmax(xu)
    packed = ops.ranges_decode(keys).cpu().numpy()
    assert np.array_equal(packed[:c.C], mn) and np.array_equal(-packed[c.C:], mx)


# ------------------------------------------------------------------------------------------ K2
def test_morph_phi(case, ops, K):
    c, x, r, d = case
    consts = K.device_constants("cuda")
    phi, dbg = ops.morph_phi(dev(r["sum"]), c.C, c.grid, consts, debug=True)
    assert np.array_equal(dbg["gray"].cpu().numpy(), d["gray"]), "gray plane"
    assert np.array_equal(dbg["gray"].cpu().numpy(), c["gray"]), "gray plane vs reference"
    edge = unpack_bits(dbg["edge_bits"], c.Wc)
    binm = unpack_bits(dbg["bin_bits"], c.Wc)
    n_edge = int((edge != d["edge"]).sum())
    n_bin = int((binm != d["binmask"]).sum())
    cnt = dbg["counts"].cpu().numpy()
    assert np.array_equal(cnt[..., 9][:, 0, 0], d["otsu_bin"]), "Otsu bin"
    assert n_bin == 0, f"{n_bin} adaptive-threshold pixels differ"
    # atan2f (<= 2 ulp) vs the oracle's correctly rounded atan2 may flip a direction bin only for
    # a pixel within an ulp of a bin boundary; none occurs in the committed cases
    assert n_edge == 0, f"{n_edge} edge pixels differ"
    assert np.array_equal(edge, c.plane_bits("edge")) and np.array_equal(binm, c.plane_bits("binmask"))
    assert np.array_equal(dbg["lbp_hist"].cpu().numpy(), d["lbp_hist"].transpose(0, 2, 3, 1))
    assert np.array_equal(cnt[..., 0], d["edge_count"])
    assert np.array_equal(cnt[..., 1], d["area"]) and np.array_equal(cnt[..., 2], d["perim"])
    assert np.array_equal(cnt[..., 3], d["euler_x4"])
    S = d["box_counts"].shape[0]
    assert np.array_equal(cnt[..., 4:4 + S], d["box_counts"].transpose(1, 2, 3, 0))
    p = phi.cpu().numpy()
    np.testing.assert_allclose(p, r["phi"], rtol=RTOL, atol=ATOL)
    np.testing.assert_allclose(p, c["phi"], rtol=RTOL, atol=ATOL)       # vs the reference
    for k in (0, 3, 4):
        assert np.array_equal(p[..., k], r["phi"][..., k])


def test_complexity(case, ops, K, W):
    c, x, r, d = case
    cm, mp, sm = packed_params(W, K)
    consts = K.device_constants("cuda")
    cpx, raw = ops.complexity(dev(r["phi"]), cm, consts, want_raw=True)
    np.testing.assert_allclose(raw.cpu().numpy(), d["complexity_raw"], rtol=RTOL, atol=ATOL)
    np.testing.assert_allclose(cpx.cpu().numpy(), r["complexity"], rtol=RTOL, atol=ATOL)
    np.testing.assert_allclose(cpx.cpu().numpy(), c["complexity"], rtol=RTOL, atol=ATOL)


def test_bit_mappers_bit_exact(case, ops, K, W):
    c, x, r, d = case
    cm, mp, sm = packed_params(W, K)
    cpx = dev(r["complexity"])
    amb = bit_ambiguous(d["bits_pre_round"])
    b = ops.bit_mapper(cpx, mp, 1.0, False).cpu().numpy()
    neq = b != r["bit_map"]
    assert int((neq & ~amb).sum()) == 0 and int(neq.sum()) <= int(amb.sum())
    assert int((b != c["bit_map_mlp"]).sum()) <= int(amb.sum())
    bc = ops.bit_mapper(cpx, mp, 1.3, True).cpu().numpy()
    np.testing.assert_allclose(bc, o.mlp_bit_mapper(r["complexity"], W["mapper"], 1.3, True), rtol=RTOL, atol=ATOL)
    d2 = {}
    lin_o = o.linear_bit_mapper(r["complexity"], 1.0, False, detail=d2)
    lin = ops.bit_mapper(cpx, None, 1.0, False).cpu().numpy()
    assert np.array_equal(lin, lin_o) and np.array_equal(lin, c["bit_map_linear"])
    assert np.all(ops.bit_mapper(cpx, mp, 10.0, False).cpu().numpy() == 8.0)   # test_smoke.py:83-84


def test_soft_mask(case, ops, K, W):
    c, x, r, d = case
    cm, mp, sm = packed_params(W, K)
    m, mt = ops.soft_mask(dev(r["bit_map"]), dev(r["abs_sum"]), c.C, sm, want_tiles=True)
    np.testing.assert_allclose(mt.cpu().numpy(), d["mask_tiles"], rtol=RTOL, atol=ATOL)
    np.testing.assert_allclose(m.cpu().numpy(), r["m"], rtol=RTOL, atol=ATOL)
    np.testing.assert_allclose(m.cpu().numpy(), c["soft_mask"], rtol=RTOL, atol=ATOL)


# ------------------------------------------------------------------------------------------ K3
def _qtable(ops, mn, mx):
    packed = dev(np.concatenate([mn, -mx]).astype(np.float32))
    return ops.build_qtable(packed)


def test_qtable_bit_exact(case, ops):
    c, x, r, d = case
    qt = _qtable(ops, r["min"], r["max"]).cpu().numpy()
    for bits in range(2, 9):
        scale, zp, _, _ = o.qparams(r["min"], r["max"], bits)
        assert np.array_equal(qt[bits - 2, :, 0], scale) and np.array_equal(qt[bits - 2, :, 1], zp)


def test_tile_quantize_codes_bit_exact(case, ops):
    """Integer codes and de-quantised values, all widths 2..8, with and without the mask."""
    c, x, r, d = case
    qt = _qtable(ops, r["min"], r["max"])
    bm = c["bit_map_rand"]
    for m_np in (None, r["m"]):
        yo, co = o.quantize_eval(x, bm, r["min"], r["max"], m_np)
        y, codes = ops.tile_quantize(dev(x), dev(bm), qt, None if m_np is None else dev(m_np), want_codes=True)
        assert np.array_equal(codes.cpu().numpy().astype(np.int16), co), "integer codes"
        assert np.array_equal(y.cpu().numpy(), yo), "de-quantised map"
    assert sha(co) == str(c["codes_rand_sha"])          # the oracle's codes are the reference's
    # MLP bit map of the case, reference's own mask -> identical to the reference output hash
    y = ops.tile_quantize(dev(x), dev(c["bit_map_mlp"]), qt, dev(c["soft_mask"]))
    assert sha(y.cpu().numpy()) == str(c["y_sha"])


def test_tile_quantize_in_place_and_reference_launcher(case, ops):
    c, x, r, d = case
    qt = _qtable(ops, r["min"], r["max"])
    bm = c["bit_map_rand"]
    yo, _ = o.quantize_eval(x, bm, r["min"], r["max"], r["m"])
    xt = dev(x)
    out = ops.tile_quantize(xt, dev(bm), qt, dev(r["m"]), out=xt)
    assert out.data_ptr() == xt.data_ptr() and np.array_equal(xt.cpu().numpy(), yo)
    # mcaq_cuda_ops.spatial_quantize signature (ops/src/mcaq_ops.cpp:22-77)
    y2 = ops.spatial_quantize(dev(x), dev(bm), dev(r["min"]).view(1, -1, 1, 1), dev(r["max"]).view(1, -1, 1, 1),
                              c.H // c.ht, c.W // c.wt, dev(r["m"]).unsqueeze(1))
    assert np.array_equal(y2.cpu().numpy(), yo)
    with pytest.raises(RuntimeError):
        ops.spatial_quantize(dev(x), dev(bm), dev(r["min"][:1]), dev(r["max"][:1]), 4, 4)


def test_tile_quantize_bf16(case, ops):
    """bf16 I/O, fp32 arithmetic: codes exact vs the oracle on the upcast input; y within one
    LSB (= scale of the assigned width) after the bf16 output rounding."""
    c, x, r, d = case
    xb = dev(x, torch.bfloat16)
    xu = xb.float().cpu().numpy()
    mn, mx = o.channel_minmax(xu)
    qt = _qtable(ops, mn, mx)
    bm = c["bit_map_rand"]
    yo, co = o.quantize_eval(xu, bm, mn, mx, r["m"])
    y, codes = ops.tile_quantize(xb, dev(bm), qt, dev(r["m"]), want_codes=True)
    assert y.dtype == torch.bfloat16
    assert np.array_equal(codes.cpu().numpy().astype(np.int16), co)
    expect = torch.from_numpy(yo).to(torch.bfloat16)      # round-to-nearest-even, like the kernel
    assert torch.equal(y.cpu(), expect)
    iy, ix = o.tile_lookup(c.H, c.W, c.ht, c.wt)
    bpix = bm[:, iy][:, :, ix].astype(int)
    lsb = np.stack([o.qparams(mn, mx, b)[0] for b in range(2, 9)])        # (7, C)
    lsb_pix = lsb[bpix - 2].transpose(0, 3, 1, 2)
    assert np.all(np.abs(y.float().cpu().numpy() - yo) <= lsb_pix)


def test_training_forward_backward(case, ops, W):
    c, x, r, d = case
    g = c.grad()
    bf = c["bit_map_frac"]
    mn, mx = c["train_run_min"], c["train_run_max"]
    qt = _qtable(ops, mn, mx)
    m_ref = c["train_soft_mask"]
    y = ops.tile_quantize_train_fwd(dev(x), dev(bf), qt, dev(m_ref))
    assert sha(y.cpu().numpy()) == str(c["train_y_sha"]), "training forward vs reference"
    dx, dbit, dm = ops.tile_quantize_train_bwd(dev(g), dev(x), dev(bf), qt, dev(m_ref))
    assert sha(dx.cpu().numpy()) == str(c["train_dx_sha"]), "dx vs reference autograd"
    dxo, dbo, dmo = o.quantize_train_bwd(g, x, bf, mn, mx, m_ref)
    np.testing.assert_allclose(dbit.cpu().numpy(), dbo, rtol=2e-3, atol=2e-3 * np.abs(dbo).max())
    np.testing.assert_allclose(dm.cpu().numpy(), dmo, rtol=2e-3, atol=2e-3 * np.abs(dmo).max())
    # mask off: d(bit_map) equals the reference's autograd (tests/test_smoke.py:103-112)
    y2 = ops.tile_quantize_train_fwd(dev(x), dev(bf), qt, None)
    assert sha(y2.cpu().numpy()) == str(c["train_nomask_y_sha"])
    _, dbit2, dm2 = ops.tile_quantize_train_bwd(dev(g), dev(x), dev(bf), qt, None)
    ref = c["train_nomask_dbit"]
    assert dm2 is None
    np.testing.assert_allclose(dbit2.cpu().numpy(), ref, rtol=2e-3, atol=2e-3 * np.abs(ref).max())


def test_ema_ranges(case, ops):
    c, x, r, d = case
    C = c.C
    s, a, keys = ops.reduce_planes(dev(x))
    packed = ops.ranges_decode(keys)
    rmin = torch.zeros(C, device="cuda")
    rmax = torch.zeros(C, device="cuda")
    ops.ranges_ema(packed, rmin, rmax, 0.99, True)
    assert np.array_equal(rmin.cpu().numpy(), c["train_run_min"]) and np.array_equal(rmax.cpu().numpy(), c["train_run_max"])
    x2 = (x * np.float32(1.5) + np.float32(0.25)).astype(np.float32)
    _, _, keys2 = ops.reduce_planes(dev(x2))
    ops.ranges_ema(ops.ranges_decode(keys2), rmin, rmax, 0.99, False)
    mn2, mx2 = o.ema_update(c["train_run_min"], c["train_run_max"], *o.channel_minmax(x2))
    assert np.array_equal(rmin.cpu().numpy(), mn2) and np.array_equal(rmax.cpu().numpy(), mx2)
    np.testing.assert_allclose(rmin.cpu().numpy(), c["ema2_min"], rtol=1e-6)
    qt = ops.build_qtable(None, rmin, rmax).cpu().numpy()
    scale, zp, _, _ = o.qparams(mn2, mx2, 5)
    assert np.array_equal(qt[3, :, 0], scale) and np.array_equal(qt[3, :, 1], zp)


# ---------------------------------------------------------------------------- odd shapes / errors
@pytest.mark.parametrize("shape", [(1, 3, 7, 9), (2, 5, 13, 21), (1, 20, 12, 12), (3, 33, 8, 16)])
def test_odd_shapes_scalar_paths(shape, ops):
    """H*W not a multiple of the vector width, C not a multiple of 16: scalar kernels."""
    from inputs import feature_map, integer_bit_map
    B, C, H, Wd = shape
    x = feature_map("noise", B, C, H, Wd, seed=77)
    s, a, keys = ops.reduce_planes(dev(x))
    so, ao = o.channel_sums(x)
    assert np.array_equal(s.cpu().numpy(), so) and np.array_equal(a.cpu().numpy(), ao)
    mn, mx = o.channel_minmax(x)
    packed = ops.ranges_decode(keys).cpu().numpy()
    assert np.array_equal(packed[:C], mn) and np.array_equal(-packed[C:], mx)
    Ht, Wt = max(1, H // 4), max(1, Wd // 4)
    bm = integer_bit_map(B, Ht, Wt, seed=5)
    qt = ops.build_qtable(dev(np.concatenate([mn, -mx])))
    yo, co = o.quantize_eval(x, bm, mn, mx, None)
    for dt in (torch.float32, torch.bfloat16):
        xt = dev(x, dt)
        if dt == torch.bfloat16:
            xu = xt.float().cpu().numpy()
            yo2, co2 = o.quantize_eval(xu, bm, mn, mx, None)
            y, codes = ops.tile_quantize(xt, dev(bm), qt, None, want_codes=True)
            assert np.array_equal(codes.cpu().numpy().astype(np.int16), co2)
            assert torch.equal(y.cpu(), torch.from_numpy(yo2).to(torch.bfloat16))
        else:
            y, codes = ops.tile_quantize(xt, dev(bm), qt, None, want_codes=True)
            assert np.array_equal(codes.cpu().numpy().astype(np.int16), co) and np.array_equal(y.cpu().numpy(), yo)


def test_errors_are_loud(ops):
    with pytest.raises(RuntimeError):
        ops.reduce_planes(torch.zeros(1, 4, 8, 8))                      # CPU tensor: no fallback
    with pytest.raises(TypeError):
        ops.reduce_planes(torch.zeros(1, 4, 8, 8, device="cuda", dtype=torch.float64))   # fp32 / bf16 / fp16 only
    x = torch.zeros(1, 4, 8, 8, device="cuda")
    with pytest.raises(RuntimeError):
        ops.tile_quantize(x, torch.zeros(2, 2, 2, device="cuda"), torch.zeros(7, 4, 2, device="cuda"))
