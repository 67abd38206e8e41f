"""Container-only: the numpy oracle against the RUNNING reference (imported read-only from /root/reference)
on cases that are NOT among the committed goldens -- other shapes, grids, temperatures, the linear mapper,
frozen calibration.  Widens the pin of tests/test_oracle_golden.py; skipped where the reference tree is
absent (the GPU box), so nothing here is needed at run time there."""
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))

import mcaq_oracle as o
from golden_util import bit_ambiguous, weights
from inputs import feature_map

RTOL, ATOL = 1e-4, 2e-6


@pytest.fixture(scope="module")
def ref():
    from ref_loader import load_reference, reference_root
    if reference_root() is None:
        pytest.skip("reference tree not mounted")
    torch.set_num_threads(min(8, os.cpu_count() or 1))
    morph, ba, qz = load_reference()
    W = weights()
    sd = lambda d: {k: torch.as_tensor(v) for k, v in d.items()}      # noqa: E731

    def build(grid):
        A = morph.MorphologicalComplexityAnalyzer(grid_size=grid, device="cpu")
        A.load_state_dict(sd(W["analyzer"]))
        Mp = ba.ComplexityToBitMappingNetwork()
        Mp.load_state_dict(sd(W["mapper"]))
        Q = qz.SpatialAdaptiveQuantization(calibration_mode="minmax", smooth_transitions=True, per_channel=True)
        Q.load_state_dict(sd(W["quantizer"]))
        return A.eval(), Mp.eval(), Q.eval(), ba.LinearBitMapper()
    return build, W


# (kind, B, C, H, W, grid, seed, temperature) -- none of these is a committed golden case
LIVE = [
    ("smooth", 1, 24, 48, 48, 8, 101, 1.0),
    ("noise", 2, 8, 28, 28, 16, 102, 1.0),
    ("smooth", 1, 20, 64, 64, 4, 103, 0.7),
    ("smooth", 2, 12, 36, 52, 8, 104, 1.3),        # rectangular, H % tile != 0 -> crop
    ("noise", 1, 32, 40, 40, 8, 105, 1.0),
    ("smooth", 1, 16, 96, 96, 8, 106, 1.0),
]


@pytest.mark.parametrize("kind,B,C,H,Wd,grid,seed,T", LIVE)
def test_hook_matches_running_reference(ref, kind, B, C, H, Wd, grid, seed, T):
    build, W = ref
    A, Mp, Q, _ = build(grid)
    x = feature_map(kind, B, C, H, Wd, seed)
    xt = torch.from_numpy(x)
    with torch.no_grad():
        c_ref = A(xt)
        b_ref = Mp(c_ref, T, return_continuous=False)
        y_ref = Q(xt, b_ref, training=False)
    d = {}
    r = o.hook_forward(x, W["analyzer"], W["mapper"], W["quantizer"], grid, T, detail=d)
    np.testing.assert_allclose(r["complexity"], c_ref.numpy(), rtol=RTOL, atol=ATOL)
    amb = bit_ambiguous(d["bits_pre_round"])
    neq = r["bit_map"] != b_ref.numpy()
    assert int((neq & ~amb).sum()) == 0, f"{int(neq.sum())} tiles differ, {int(amb.sum())} ambiguous"
    if not neq.any():
        np.testing.assert_allclose(r["y"], y_ref.numpy(), rtol=RTOL, atol=ATOL)
        # the integer codes, recomputed the reference's way from ITS ranges, equal the oracle's
        mn, mx = xt.amin(dim=(0, 2, 3)).numpy(), xt.amax(dim=(0, 2, 3)).numpy()
        assert np.array_equal(mn, r["min"]) and np.array_equal(mx, r["max"])


@pytest.mark.parametrize("kind,B,C,H,Wd,grid,seed,T", LIVE[:3])
def test_linear_mapper_and_frozen_ranges_match_running_reference(ref, kind, B, C, H, Wd, grid, seed, T):
    build, W = ref
    A, _, Q, Lin = build(grid)
    x = feature_map(kind, B, C, H, Wd, seed)
    xt = torch.from_numpy(x)
    with torch.no_grad():
        c_ref = A(xt)
        b_ref = Lin(c_ref, T)
        # frozen calibration: statistics of ANOTHER batch, then quantise this one (quantization.py:647-649)
        x0 = torch.from_numpy(feature_map("noise", B, C, H, Wd, seed + 50))
        Q.train()
        Q(x0, b_ref, training=True)
        Q.freeze_calibration()
        Q.eval()
        y_ref = Q(xt, b_ref, training=False)
        rmin, rmax = Q.running_min.reshape(-1).numpy(), Q.running_max.reshape(-1).numpy()
    r = o.hook_forward(x, W["analyzer"], None, W["quantizer"], grid, T, frozen_minmax=(rmin, rmax))
    assert np.array_equal(r["bit_map"], b_ref.numpy()), "linear (quantile) mapper"
    np.testing.assert_allclose(r["y"], y_ref.numpy(), rtol=RTOL, atol=ATOL)
    # first EMA step = the batch statistics (quantization.py:340-347)
    mn0, mx0 = o.channel_minmax(x0.numpy())
    assert np.array_equal(rmin, mn0) and np.array_equal(rmax, mx0)


def test_score_image_matches_running_reference(ref):
    build, W = ref
    A, _, _, _ = build(8)
    x = feature_map("smooth", 2, 3, 160, 160, 207)
    with torch.no_grad():
        s_ref = A.score_image(torch.from_numpy(x)).numpy()
    np.testing.assert_allclose(o.score_image(x, W["analyzer"]["feature_weights"], 8), s_ref, rtol=RTOL, atol=ATOL)


@pytest.mark.parametrize("kind,B,C,H,Wd,grid,seed,T", LIVE[:4])
def test_training_forward_backward_match_running_reference(ref, kind, B, C, H, Wd, grid, seed, T):
    """Fractional-bit compose and its autograd (quantization.py:699-727, 69-118) on fresh cases: forward and
    dx bit-identical, d(bit_map) within the fp32 summation-order tolerance."""
    build, W = ref
    _, _, Q, _ = build(grid)
    Q.smooth_transitions = False                   # mask off: d(bit_map) is the pure tile-sum formula
    x = feature_map(kind, B, C, H, Wd, seed)
    g = feature_map("noise", B, C, H, Wd, seed + 500)
    tile = o.tile_size(H, grid)
    ht, wt = H // tile, Wd // tile
    rng = np.random.default_rng(seed)
    bf = (rng.random((B, ht, wt), dtype=np.float32) * 6.2 + 1.9).astype(np.float32)
    xt = torch.from_numpy(x).requires_grad_(True)
    bt = torch.from_numpy(bf).requires_grad_(True)
    Q.train()
    y = Q(xt, bt, training=True)
    y.backward(torch.from_numpy(g))
    mn, mx = Q.running_min.reshape(-1).numpy(), Q.running_max.reshape(-1).numpy()
    yo = o.quantize_train_fwd(x, bf, mn, mx, None)[0]
    dxo, dbo, _ = o.quantize_train_bwd(g, x, bf, mn, mx, None)
    assert np.array_equal(y.detach().numpy(), yo), "training forward"
    np.testing.assert_allclose(xt.grad.numpy(), dxo, rtol=1e-6, atol=1e-7)
    ref_db = bt.grad.numpy()
    np.testing.assert_allclose(dbo, ref_db, rtol=2e-3, atol=2e-3 * np.abs(ref_db).max())
