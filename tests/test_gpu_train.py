"""GPU parity of the training forms of K3: the vector kernels against the scalar kernels (y and dx
bit for bit), and the distillation-fused forward / backward (train.py:599-610) against
the same loss composed from the plain entry points with torch's own mse_loss autograd.
The reference parity of the training forward / dx itself (sha256 of the real reference's outputs) is
tests/test_gpu_parity.py::test_training_forward_backward, which runs through the same dispatcher."""
import numpy as np
import pytest
import torch

from golden_util import Case, weights
from inputs import feature_map

pytestmark = pytest.mark.gpu

VEC_CASES = ["c3_v8n_smooth", "c4_v8n_smooth", "c5_v8n_smooth", "c3_grid4", "c3_grid16", "parity_2x8x32"]


@pytest.fixture(scope="module")
def ops():
    from mcaq_yolo_b200 import ops as _ops
    return _ops


@pytest.fixture(scope="module")
def lib():
    from mcaq_yolo_b200 import _lib
    return _lib.load()


def dev(a, dtype=None):
    t = torch.from_numpy(np.ascontiguousarray(a)).cuda()
    return t if dtype is None else t.to(dtype)


def _setup(name, ops, dtype):
    c = Case(name)
    x = dev(c.x(), dtype)
    g = dev(c.grad(), dtype)
    bf = dev(c["bit_map_frac"])
    qt = ops.build_qtable(None, dev(c["train_run_min"]).reshape(-1), dev(c["train_run_max"]).reshape(-1))
    m = dev(c["train_soft_mask"]).reshape(c.B, c.H, c.W)
    return c, x, g, bf, qt, m


@pytest.mark.parametrize("name", VEC_CASES)
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("with_mask", [True, False])
def test_vector_kernels_equal_scalar_kernels(name, dtype, with_mask, ops, lib):
    c, x, g, bf, qt, m = _setup(name, ops, dtype)
    assert ops.kd_geometry_ok(x, bf), "case is meant to exercise the vector path"
    mm = m if with_mask else None
    try:
        lib.mcaq_debug_train_scalar(1)
        y_s = ops.tile_quantize_train_fwd(x, bf, qt, mm)
        dx_s, db_s, dm_s = ops.tile_quantize_train_bwd(g, x, bf, qt, mm)
    finally:
        lib.mcaq_debug_train_scalar(0)
    y_v = ops.tile_quantize_train_fwd(x, bf, qt, mm)
    dx_v, db_v, dm_v = ops.tile_quantize_train_bwd(g, x, bf, qt, mm)
    assert torch.equal(y_v, y_s), "training forward: vector vs scalar kernel"
    assert torch.equal(dx_v, dx_s), "dx: vector vs scalar kernel"
    scale = float(db_s.abs().max()) + 1e-12
    np.testing.assert_allclose(db_v.cpu().numpy(), db_s.cpu().numpy(), rtol=2e-3, atol=2e-3 * scale)
    if with_mask:
        np.testing.assert_allclose(dm_v.cpu().numpy(), dm_s.cpu().numpy(), rtol=1e-4,
                                   atol=1e-4 * float(dm_s.abs().max()))
    else:
        assert dm_v is None


def test_integer_bits_reduce_to_inference_codes(ops):
    """With an integral bit map the fractional compose is exactly the eval quantiser."""
    c, x, g, bf, qt, m = _setup("c3_v8n_smooth", ops, torch.float32)
    bi = dev(c["bit_map_mlp"])
    y_t = ops.tile_quantize_train_fwd(x, bi, qt, m)
    y_e = ops.tile_quantize(x, bi, qt, m)
    assert torch.equal(y_t, y_e)


@pytest.mark.parametrize("name", ["c3_v8n_smooth", "c5_v8n_smooth", "parity_2x8x32"])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("with_mask", [True, False])
def test_distillation_fused_forward_backward(name, dtype, with_mask, ops):
    c, x, g, bf, qt, m = _setup(name, ops, dtype)
    mm = m if with_mask else None
    teacher = dev(feature_map("smooth", c.B, c.C, c.H, c.W, c.seed + 77)) * 0.9 + 0.05
    y0 = ops.tile_quantize_train_fwd(x, bf, qt, mm)
    y, kd_sum = ops.tile_quantize_train_fwd_kd(x, bf, qt, mm, teacher)
    assert torch.equal(y, y0), "the fused forward must not change y"
    mse = torch.nn.functional.mse_loss(y0.double(), teacher.double())
    np.testing.assert_allclose(float(kd_sum) / y.numel(), float(mse), rtol=2e-6)
    # backward: g_t = g + coef (y - t)
    coef = torch.tensor([0.37 * 2.0 / y.numel() * 1000.0], device="cuda")     # large enough to matter
    dx, db, dm = ops.tile_quantize_train_bwd_kd(g, x, bf, qt, mm, teacher, coef)
    gt = g.float() + coef * (y0.float() - teacher)
    if dtype == torch.float32:
        dx_r, db_r, dm_r = ops.tile_quantize_train_bwd(gt, x, bf, qt, mm)
        assert torch.equal(dx, dx_r), "fp32: same operation order as the composed path"
        tol = 2e-3
    else:
        # composed path in fp32 on the upcast tensors (the fused kernel keeps g_t in fp32)
        dx_r, db_r, dm_r = ops.tile_quantize_train_bwd(gt, x.float(), bf, qt, mm)
        np.testing.assert_allclose(dx.float().cpu().numpy(), dx_r.cpu().numpy(), rtol=8e-3,
                                   atol=8e-3 * float(dx_r.abs().max()))
        tol = 1e-2
    np.testing.assert_allclose(db.cpu().numpy(), db_r.cpu().numpy(), rtol=tol, atol=tol * float(db_r.abs().max()))
    if with_mask:
        np.testing.assert_allclose(dm.cpu().numpy(), dm_r.cpu().numpy(), rtol=tol, atol=tol * float(dm_r.abs().max()))


def test_kd_entry_points_reject_scalar_geometry(ops):
    x = torch.randn(1, 3, 7, 9, device="cuda")
    bm = torch.full((1, 2, 2), 4.5, device="cuda")
    qt = ops.build_qtable(None, torch.full((3,), -3.0, device="cuda"), torch.full((3,), 3.0, device="cuda"))
    assert not ops.kd_geometry_ok(x, bm)
    with pytest.raises(RuntimeError, match="vector path"):
        ops.tile_quantize_train_fwd_kd(x, bm, qt, None, torch.zeros_like(x))


@pytest.mark.parametrize("shape", [(1, 3, 7, 9), (2, 8, 50, 50)])
def test_module_kd_composed_fallback_for_odd_geometry(shape):
    from mcaq_yolo_b200 import modules as M
    W = weights()
    _, _, q = M.build_fixture_modules(W, "cuda")
    q.train()
    torch.manual_seed(3)
    x = torch.randn(*shape, device="cuda", requires_grad=True)
    bm = torch.rand(shape[0], 2, 2, device="cuda") * 6 + 2
    q.kd_teacher = torch.randn(*shape, device="cuda")
    y = q(x, bm, training=True)
    assert q.kd_feature_loss is not None
    np.testing.assert_allclose(float(q.kd_feature_loss), float(torch.nn.functional.mse_loss(y, q.kd_teacher)), rtol=1e-6)


def test_module_level_distillation_matches_unfused_autograd():
    """SpatialAdaptiveQuantization with kd_teacher set: same y, same loss and the same gradients for
    x, the bit map and the soft-mask net as the unfused composition (quantiser + F.mse_loss)."""
    from mcaq_yolo_b200 import modules as M
    W = weights()
    c = Case("small_smooth")
    g = dev(c.grad())
    teacher = dev(feature_map("smooth", c.B, c.C, c.H, c.W, c.seed + 77))

    def run(fused):
        _, _, q = M.build_fixture_modules(W, "cuda")
        q.train()
        x = dev(c.x()).requires_grad_(True)
        bf = dev(c["bit_map_frac"]).requires_grad_(True)
        q.kd_teacher = teacher if fused else None
        y = q(x, bf, training=True)
        kd = q.kd_feature_loss if fused else torch.nn.functional.mse_loss(y.float(), teacher)
        loss = (y * g).sum() + 250.0 * kd
        loss.backward()
        return (y.detach(), kd.detach(), x.grad, bf.grad, q.soft_mask.net[0].weight.grad, q.soft_mask.net[2].bias.grad)

    f, u = run(True), run(False)
    assert ops_geometry_used(c)
    assert torch.equal(f[0], u[0])
    np.testing.assert_allclose(float(f[1]), float(u[1]), rtol=2e-6)
    np.testing.assert_allclose(f[2].cpu().numpy(), u[2].cpu().numpy(), rtol=1e-5, atol=1e-6 * float(u[2].abs().max()))
    for a, b in zip(f[3:], u[3:]):
        np.testing.assert_allclose(a.cpu().numpy(), b.cpu().numpy(), rtol=5e-3, atol=5e-3 * float(b.abs().max()))


def ops_geometry_used(c):
    from mcaq_yolo_b200 import ops as _ops
    return _ops.kd_geometry_ok(torch.empty(c.B, c.C, c.H, c.W), torch.empty(c.B, c.ht, c.wt))


@pytest.mark.parametrize("shape", [(64, 80, 80), (128, 40, 40), (256, 20, 20)])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_training_full_size_vector_equals_scalar(shape, dtype, ops, lib):
    """BASELINE configs[3] per-GPU size (16 images): vector == scalar kernels on y and dx, and the
    tile sums of d(bit_map) match an fp64 torch restatement."""
    C, H, Wd = shape
    B = 16
    gen = torch.Generator(device="cuda")
    gen.manual_seed(11)
    x = (torch.randn(B, C, H, Wd, device="cuda", generator=gen) * 1.7 + 0.3).to(dtype)
    g = torch.randn(B, C, H, Wd, device="cuda", generator=gen).to(dtype)
    tile = ops.tile_size(H, 8)
    ht = H // tile
    bf = torch.rand(B, ht, ht, device="cuda", generator=gen) * 6.5 + 1.8
    m = torch.rand(B, H, Wd, device="cuda", generator=gen) * 0.2 + 0.8
    mn = x.float().amin(dim=(0, 2, 3)) * 0.9
    mx = x.float().amax(dim=(0, 2, 3)) * 0.9            # some values clip
    qt = ops.build_qtable(None, mn, mx)
    try:
        lib.mcaq_debug_train_scalar(1)
        y_s = ops.tile_quantize_train_fwd(x, bf, qt, m)
        dx_s, db_s, dm_s = ops.tile_quantize_train_bwd(g, x, bf, qt, m)
    finally:
        lib.mcaq_debug_train_scalar(0)
    y_v = ops.tile_quantize_train_fwd(x, bf, qt, m)
    dx_v, db_v, dm_v = ops.tile_quantize_train_bwd(g, x, bf, qt, m)
    assert torch.equal(y_v, y_s) and torch.equal(dx_v, dx_s)
    # fp64 restatement of d(bit_map) = sum_tile g m (Q_hi - Q_lo) from two integer-bit forwards
    lo = bf.floor().clamp(2, 8)
    hi = (lo + 1).clamp(max=8)
    q_lo = ops.tile_quantize(x.float(), lo, qt, None).double()
    q_hi = ops.tile_quantize(x.float(), hi, qt, None).double()
    contrib = (g.double() * m.double().unsqueeze(1) * (q_hi - q_lo)).sum(1)
    ref = contrib.reshape(B, ht, tile, ht, tile).sum(dim=(2, 4))
    np.testing.assert_allclose(db_v.cpu().numpy(), ref.cpu().numpy(), rtol=2e-3, atol=2e-3 * float(ref.abs().max()))
    np.testing.assert_allclose(dm_v.cpu().numpy(), dm_s.cpu().numpy(), rtol=2e-4, atol=2e-4 * float(dm_s.abs().max()))
