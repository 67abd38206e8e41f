"""tools/backend_agreement.py (the surrogate-vs-cv2 diagnostic of the reference's
scripts/backend_agreement.py:47-102): the host-side cv2 recipes on known shapes, against the RUNNING reference's own
cv2 backend (container only), and one end-to-end pass with the native analyzer (GPU)."""
import math
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
cv2 = pytest.importorskip("cv2")
import backend_agreement as BA  # noqa: E402


def test_uniform_lbp_codes_on_known_patterns():
    flat = np.full((12, 12), 77, np.uint8)
    codes = BA.lbp_uniform_8_1(flat)
    assert np.all(codes[1:-1, 1:-1] == 8)                  # every sample equals the centre: eight ones, no transition
    assert codes[0, 0] == 3                                # corner: the three in-image samples (E, SE, S) are >= centre
    peak = flat.copy(); peak[6, 6] = 200
    assert BA.lbp_uniform_8_1(peak)[6, 6] == 0             # a local maximum: no sample reaches the centre
    stripes = np.tile(np.array([0, 255], np.uint8), (12, 6))
    c = BA.lbp_uniform_8_1(stripes)
    assert set(np.unique(c[2:-2, 2:-2])) <= {2.0, 8.0, 9.0}   # bright columns: all >=; dark ones: two opposite arcs
    assert 0.0 < BA.texture_entropy(stripes) < 1.0 and BA.texture_entropy(flat) < 0.35


def test_cv2_recipes_on_a_disc_and_on_noise():
    disc = np.zeros((64, 64), np.uint8)
    cv2.circle(disc, (32, 32), 20, 255, -1)
    assert BA.contour_complexity(disc) < 0.25              # near-circular: P^2 / (4 pi A) close to 1
    assert 0.01 < BA.edge_density(disc) < 0.08             # one thin ring of ~126 edge pixels in 4096
    rng = np.random.default_rng(0)
    noise = rng.integers(0, 256, (64, 64)).astype(np.uint8)
    assert BA.edge_density(noise) > BA.edge_density(disc)
    assert BA.gradient_variance(noise) > BA.gradient_variance(disc) > 0.0
    e = (BA.canny_otsu(noise) > 0).astype(np.uint8)
    assert 1.0 <= BA.fractal_dimension(e) <= 2.0 and BA.fractal_dimension(np.zeros((64, 64), np.uint8)) == 1.0
    assert BA.tile_size(640) == 64 and BA.tile_size(80) == 8 and BA.tile_size(20) == 4
    imgs = BA.synthetic_images(1, 128, 3)
    phi, det = BA.cv2_phi_tiles(imgs, 8)
    assert phi.shape == (1, 8, 8, 8) and np.isfinite(phi).all() and phi.min() >= 0.0 and phi.max() <= 1.0
    np.testing.assert_allclose(phi[..., 5], phi[..., 0] * phi[..., 1], rtol=1e-6)
    np.testing.assert_allclose(phi[..., 7], np.sqrt(phi[..., 3] * phi[..., 4]), rtol=1e-6, atol=1e-7)
    a, b = np.arange(10.0), np.arange(10.0) ** 3
    assert BA.spearman(a, b) == pytest.approx(1.0) and BA.pearson(a, b) < 1.0 and math.isnan(BA.pearson(a, np.ones(10)))


def test_cv2_recipes_match_the_running_reference_backend():
    """Container only: the reference's `metric_backend='cv2'` path (morphology.py:741-797) with skimage's LBP replaced
    by the restatement here (skimage is not in the image): the other four metrics and the three products are its own."""
    from ref_loader import load_reference, reference_root
    if reference_root() is None:
        pytest.skip("reference tree not mounted")
    import torch
    morph, _, _ = load_reference()
    morph.local_binary_pattern = lambda gray, P, R, method: BA.lbp_uniform_8_1(gray)
    ana = morph.MorphologicalComplexityAnalyzer(grid_size=4, device="cpu", metric_backend="cv2").eval()
    imgs = BA.synthetic_images(2, 128, 11)
    phi_ref, det_ref = ana.compute_phi_tiles(torch.from_numpy(imgs))
    phi, det = BA.cv2_phi_tiles(imgs, 4)
    assert tuple(phi_ref.shape) == phi.shape == (2, 4, 4, 8)
    np.testing.assert_allclose(phi, phi_ref.numpy(), rtol=1e-6, atol=1e-7)
    for m in BA.METRICS:
        np.testing.assert_allclose(det[m], det_ref[m].numpy(), rtol=1e-6, atol=1e-7)


@pytest.mark.gpu
def test_agreement_table_with_the_native_analyzer():
    imgs = BA.synthetic_images(3, 256, 5)
    res = BA.run(imgs, grid=8, batch=2)
    assert res["_tiles_per_image"] == 64
    for m in BA.METRICS + ["fused_C"]:
        r = res[m]
        assert 0.0 <= r["mean_gpu"] <= 1.0 and 0.0 <= r["mean_cv2"] <= 1.0
        assert math.isnan(r["pearson"]) or -1.0 <= r["pearson"] <= 1.0
    # the two backends measure the same things: edge density and gradient variance correlate strongly on these scenes
    assert res["edge"]["spearman"] > 0.5 and res["gradient"]["spearman"] > 0.5
