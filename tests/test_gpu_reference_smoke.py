"""The reference's own smoke suite (mcaq_yolo/tests/test_smoke.py), restated against the
native modules on the GPU: same constructions, same assertions, freshly initialised (not fixture) weights.
test_cuda_kernel_parity lives in test_gpu_modules.py::test_reference_cuda_parity_test_shape, the Euler
known answers in test_oracle_golden.py + the bit-exact count comparison of test_gpu_parity.py."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def M():
    from mcaq_yolo_b200 import modules
    return modules


@pytest.mark.parametrize("H", [640, 160, 80, 40, 20])             # the reference's own list starts at 640
def test_phi_tiles_shapes(H, M):                                     # test_smoke.py:33-47
    a = M.MorphologicalComplexityAnalyzer(device="cuda")
    torch.manual_seed(0)
    x = torch.rand(2, 3, H, H, device="cuda")
    phi, detailed = a.compute_phi_tiles(x)
    tile = a._tile_size(H)
    ht = H // tile
    assert phi.shape == (2, ht, ht, 8), phi.shape
    assert set(detailed) == {"fractal", "texture", "gradient", "edge", "contour"}
    assert tile >= 4 and (tile & (tile - 1)) == 0
    assert float(phi.min()) >= 0.0 and float(phi.max()) <= 1.0 + 1e-5


def test_phi_tiles_on_image_sized_planes(M):
    """Planes beyond the fused kernel's on-chip budget (> 160 columns or tiles > 32 pixels: raw 640 / 1280
    images, utils/dataset.py:345-353) take the plane pipeline (csrc/morph_planes.cu), same results contract."""
    from mcaq_yolo_b200 import ops
    a = M.MorphologicalComplexityAnalyzer(device="cuda")
    assert not ops.morph_fits(1, 3, 640, 640, 8) and not ops.morph_fits(1, 3, 192, 192, 8)
    assert ops.morph_fits(64, 64, 80, 80, 8) and ops.morph_fits(8, 128, 160, 160, 8)
    for H, ht in ((1280, 10), (192, 12), (200, 12)):
        phi, _ = a.compute_phi_tiles(torch.rand(1, 3, H, H, device="cuda"))
        assert phi.shape == (1, ht, ht, 8) and bool(torch.isfinite(phi).all())
        assert float(phi.min()) >= 0.0 and float(phi.max()) <= 1.0 + 1e-5
    s = a.score_image(torch.rand(3, 3, 640, 640, device="cuda"))
    assert s.shape == (3,) and float(s.min()) >= 0.0 and float(s.max()) <= 1.0


def test_analyzer_forward_range_and_grad(M):                         # test_smoke.py:50-59
    a = M.MorphologicalComplexityAnalyzer(device="cuda")
    a.train()
    x = torch.rand(2, 16, 80, 80, device="cuda")
    c = a(x)
    assert c.dim() == 3 and 0.0 <= float(c.min()) and float(c.max()) <= 1.0
    c.sum().backward()
    grads = [p.grad for p in a.complexity_mlp.parameters() if p.grad is not None]
    assert grads and any(float(g.abs().sum()) > 0 for g in grads)


def test_score_image_deterministic(M):                               # test_smoke.py:62-67
    a = M.MorphologicalComplexityAnalyzer(device="cuda")
    x = torch.rand(1, 3, 160, 160, device="cuda")
    s1, s2 = a.score_image(x), a.score_image(x)
    assert torch.equal(s1, s2)
    assert 0.0 <= float(s1) <= 1.0


def test_bit_mapper_range_and_temperature(M):                        # test_smoke.py:74-84
    m = M.ComplexityToBitMappingNetwork(min_bits=2, max_bits=8).cuda()
    m.eval()
    c = torch.rand(2, 8, 8, device="cuda")
    b = m(c, temperature=1.0)
    assert b.shape == (2, 8, 8)
    assert float(b.min()) >= 2.0 and float(b.max()) <= 8.0
    assert torch.equal(b, torch.round(b))
    b10 = m(c, temperature=10.0)
    assert torch.equal(b10, torch.full_like(b10, 8.0))


def test_bit_mapper_gradient_through_clamp_and_round(M):             # test_smoke.py:87-96
    m = M.ComplexityToBitMappingNetwork(min_bits=2, max_bits=8).cuda()
    m.train()
    c = torch.rand(2, 8, 8, device="cuda")
    b = m(c, temperature=10.0)
    (b.mean() - 4.0).pow(2).backward()
    grads = [p.grad for p in m.mapping_network.parameters() if p.grad is not None]
    assert grads and any(float(g.abs().sum()) > 0 for g in grads), "clamp killed the gradient"


def test_fractional_bit_gradient_to_bit_map(M):                      # test_smoke.py:103-112
    q = M.SpatialAdaptiveQuantization(smooth_transitions=False).cuda()
    q.train()
    x = torch.randn(1, 4, 16, 16, device="cuda")
    bit_map = torch.full((1, 4, 4), 4.5, device="cuda", requires_grad=True)
    y = q(x, bit_map, training=True)
    assert y.shape == x.shape
    y.pow(2).mean().backward()
    assert bit_map.grad is not None and float(bit_map.grad.abs().sum()) > 0


def test_learned_soft_mask_near_identity_init(M):                    # test_smoke.py:115-126
    mask = M.LearnedSoftMask().cuda()
    x = torch.randn(2, 8, 32, 32, device="cuda")
    bit_map = torch.full((2, 4, 4), 4.0, device="cuda")
    m = mask(bit_map, x)
    assert m.shape == (2, 1, 32, 32)
    assert float(m.min()) > 0.9, float(m.min())
    m.sum().backward()
    g_first = mask.net[0].weight.grad
    assert g_first is not None and float(g_first.abs().sum()) > 0
    with torch.no_grad():                                             # the kernel path (no grad) gives the same mask
        m2 = mask(bit_map, x)
    assert m2.shape == (2, 1, 32, 32) and torch.allclose(m2, m.detach(), rtol=1e-5, atol=1e-6)


def test_calibration_freeze(M):                                      # test_smoke.py:129-139
    q = M.SpatialAdaptiveQuantization(smooth_transitions=False).cuda()
    q.train()
    x = torch.randn(2, 4, 16, 16, device="cuda")
    bit_map = torch.full((2, 4, 4), 4.0, device="cuda")
    _ = q(x, bit_map, training=True)
    assert q.running_min is not None
    frozen_min = q.running_min.clone()
    q.freeze_calibration()
    _ = q(torch.randn(2, 4, 16, 16, device="cuda") * 100, bit_map, training=True)
    assert torch.equal(q.running_min, frozen_min), "stats moved after freeze"


def test_linear_bit_mapper_spatial_variance(M):                      # test_smoke.py:188-196
    m = M.LinearBitMapper(min_bits=2, max_bits=8)
    c = (torch.linspace(0, 1, 16).reshape(1, 4, 4) * 0.05 + 0.4).cuda()    # narrow absolute range
    b = m(c, temperature=1.0)
    assert float(b.min()) == 2.0 and float(b.max()) == 8.0              # normalisation spreads it
    assert torch.unique(b).numel() >= 5


def test_linear_bit_mapper_flat_map_absolute_fallback(M):            # test_smoke.py:199-211
    m = M.LinearBitMapper()
    for value, bits in ((0.5, 5.0), (0.0, 2.0), (1.0, 8.0)):
        b = m(torch.full((1, 8, 8), value, device="cuda"))
        assert torch.all(b == bits), (value, float(b.min()), float(b.max()))
