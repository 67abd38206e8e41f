"""Round-2 GPU parity cases (through the C ABI, against the oracle and tests/golden/r2/*.npz written by
tools/make_golden_r2.py from the RUNNING reference):

* one full BASELINE configs[1] batch (64 images, bf16, C3/C4/C5) against the oracle on the upcast input:
  bit maps exact, integer codes exact, y within 1 LSB of the assigned width;
* the calibration pass (eval-mode quantizer called with training=True, models/mcaq_yolo.py:446);
* the hook with normalize_complexity=True (models/mcaq_yolo.py:427-432);
* a non-monotone MLP mapper (no step table) and a constrained mapper whose weights violate Eq.18;
* fp16 feature maps (what the hooks see under torch.autocast, train.py:192, 582, 748);
* the ambiguity set is EMPTY on every committed case (so "exact outside the set" means exact).
"""
import os

import numpy as np
import pytest
import torch

import mcaq_oracle as o
from golden_util import CASE_NAMES, GOLDEN_DIR, Case, bit_ambiguous, sha, weights
from inputs import feature_map

pytestmark = pytest.mark.gpu
RTOL, ATOL = 1e-4, 2e-6


def r2(name):
    return np.load(os.path.join(GOLDEN_DIR, "r2", name + ".npz"))


def sd(d):
    return {k: torch.as_tensor(v) for k, v in d.items()}


@pytest.fixture(scope="module")
def M():
    from mcaq_yolo_b200 import modules
    return modules


@pytest.fixture(scope="module")
def W():
    return weights()


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


# ------------------------------------------------------------------ full configs[1] batch vs the oracle
def test_full_batch64_bf16_against_oracle(M, W):
    """BASELINE configs[1] itself (not a property): 64 images x (64x80x80, 128x40x40, 256x20x20), bf16,
    the fused three-launch hook against the oracle on the fp32-upcast maps."""
    from mcaq_yolo_b200 import constants as K, fused, ops
    a, m, _ = M.build_fixture_modules(W, "cuda")
    for si, (C, H) in enumerate(((64, 80), (128, 40), (256, 20))):
        q = M.build_fixture_modules(W, "cuda")[2]
        xb = torch.from_numpy(feature_map("smooth", 64, C, H, H, 70 + si)).cuda().to(torch.bfloat16)
        with torch.no_grad():
            rec, _ = fused.fused_scale_forward(xb, a, m, q, 1.0, None, layer=si)
        xu = xb.float().cpu().numpy()
        d = {}
        r = o.hook_forward(xu, W["analyzer"], W["mapper"], W["quantizer"], 8, 1.0, detail=d)
        # 6400 tiles per scale: a tile whose continuous bit value sits within 1e-4 of a .5 boundary does occur
        # here (unlike in the small committed cases); bit maps must be EQUAL everywhere else
        amb = bit_ambiguous(d["bits_pre_round"])
        assert int(amb.sum()) <= 4, f"{int(amb.sum())} ambiguous tiles"
        bm = rec["bit_map"].cpu().numpy()
        nd = int(((bm != r["bit_map"]) & ~amb).sum())
        assert nd == 0, f"C{3 + si}: {nd} bit-map tiles differ outside the ambiguity set"
        if not np.array_equal(bm, r["bit_map"]):          # continue on the GPU's choice for the ambiguous tile(s)
            r["m"] = o.soft_mask(bm, r["abs_sum"], C, W["quantizer"])
            r["y"], r["codes"] = o.quantize_eval(xu, bm, r["min"], r["max"], r["m"])
        np.testing.assert_allclose(rec["complexity"].cpu().numpy(), r["complexity"], rtol=RTOL, atol=ATOL)
        # integer codes: K3 on the same bit map / ranges / mask the fused launch used
        with torch.no_grad():
            qt = ops.build_qtable(None, dev(r["min"]), dev(r["max"]))
            mask = ops.soft_mask(rec["bit_map"], ops.reduce_planes(xb, want_ranges=False)[1], C,
                                 K.pack_soft_mask(q.soft_mask))
            y2, codes = ops.tile_quantize(xb, rec["bit_map"], qt, mask, want_codes=True)
        assert np.array_equal(codes.cpu().numpy().astype(np.int16), r["codes"].astype(np.int16)), "integer codes"
        assert torch.equal(y2, rec["features_q"]), "fused K3 and table K3 disagree"
        # y: the bf16 rounding of (code - zp) * scale * m; 1 LSB of the assigned width is the bar
        y = rec["features_q"].float().cpu().numpy()
        ty, tx = o.tile_lookup(H, H, bm.shape[1], bm.shape[2])
        bits_px = bm[:, ty][:, :, tx]                                        # (B, H, W)
        scale = np.maximum(r["max"] - r["min"], 1e-8)[None, :, None, None] / (2.0 ** bits_px[:, None] - 1.0)
        lsb = scale * np.abs(r["m"])[:, None] + np.abs(r["y"]) * 2.0 ** -8   # + one bf16 ulp of the value
        assert np.all(np.abs(y - r["y"]) <= lsb), "y differs from the oracle by more than 1 LSB"
        exact = torch.equal(rec["features_q"].cpu(), torch.from_numpy(r["y"]).to(torch.bfloat16))
        frac = float((rec["features_q"].cpu() == torch.from_numpy(r["y"]).to(torch.bfloat16)).float().mean())
        assert exact or frac > 0.999, f"only {frac:.5f} of y equals the bf16-rounded oracle value"


# ------------------------------------------------------------------ calibration pass
def test_calibration_pass_eval_module_training_arg(M, W):
    """MCAQYOLO.calibrate(): the model is in eval mode and the hook passes training=True.  The reference
    then updates the EMA but quantises with the CURRENT batch's min / max (quantization.py:415-417);
    from the second batch on that differs from quantising with the EMA."""
    g = r2("calib_c4")
    B, C, H = [int(v) for v in g["cfg"][:3]]
    _, _, q = M.build_fixture_modules(W, "cuda")
    q.eval()
    bm = dev(g["bit_map"])
    xs = [torch.from_numpy(feature_map("smooth", B, C, H, H, 31 + i)) * (1.0 + 0.5 * i) + 0.25 * i for i in range(3)]
    with torch.no_grad():
        for i in range(2):
            y = q(xs[i].cuda(), bm, training=True)
            assert np.array_equal(q.running_min.cpu().numpy().ravel(), g[f"run_min{i}"]), f"EMA min after batch {i}"
            assert np.array_equal(q.running_max.cpu().numpy().ravel(), g[f"run_max{i}"]), f"EMA max after batch {i}"
            np.testing.assert_allclose(y.cpu().numpy()[:, ::3, ::5, ::7], g[f"y{i}_sub"], rtol=RTOL, atol=ATOL,
                                       err_msg=f"calibration-pass output of batch {i}")
        q.freeze_calibration()
        y = q(xs[2].cuda(), bm, training=False)
        np.testing.assert_allclose(y.cpu().numpy()[:, ::3, ::5, ::7], g["y2_sub"], rtol=RTOL, atol=ATOL)
    # and a TRAIN-mode module does use the EMA (quantization.py:415-417): batch 1 must then differ
    _, _, q2 = M.build_fixture_modules(W, "cuda")
    q2.train()
    with torch.no_grad():
        q2(xs[0].cuda(), bm, training=True)
        y_tr = q2(xs[1].cuda(), bm, training=True)
    assert not np.allclose(y_tr.cpu().numpy()[:, ::3, ::5, ::7], g["y1_sub"], rtol=1e-3, atol=1e-3)


# ------------------------------------------------------------------ normalize_complexity
@pytest.mark.parametrize("linear", [False, True])
def test_hook_with_normalize_complexity(linear, M, W):
    g = r2("normalize_c3")
    B, C, H = [int(v) for v in g["cfg"][:3]]
    x = torch.from_numpy(feature_map("smooth", B, C, H, H, 41)).cuda()
    a, m, q = M.build_fixture_modules(W, "cuda", linear_mapper=linear)
    with torch.no_grad():
        rec = M.mcaq_hook_forward(x, a, m, q, temperature=1.0, normalize_complexity=True)
    np.testing.assert_allclose(rec["complexity"].cpu().numpy(), g["complexity_norm"], rtol=RTOL, atol=5e-6)
    ref = g["bit_map_linear" if linear else "bit_map_mlp"]
    bm = rec["bit_map"].cpu().numpy()
    # tiles whose normalised complexity sits on a rounding boundary of the mapper are excluded (none expected)
    assert int((bm != ref).sum()) <= 1, f"{(bm != ref).sum()} tiles differ"
    if not linear and np.array_equal(bm, ref):
        np.testing.assert_allclose(rec["features_q"].cpu().numpy()[:, ::3, ::5, ::7], g["y_sub"], rtol=RTOL, atol=ATOL)


def test_fused_hook_object_with_normalize_flag(M, W):
    """FusedMcaqHook takes the module-by-module path when the model sets normalize_complexity."""
    from mcaq_yolo_b200.fused import FusedMcaqHook
    g = r2("normalize_c3")
    B, C, H = [int(v) for v in g["cfg"][:3]]
    a, m, q = M.build_fixture_modules(W, "cuda")

    class Model:
        training = False
        normalize_complexity = True
        complexity_analyzer, bit_mapper = a, m
        quantizers = {"4": q}
        _mcaq_state = {"active": True, "temperature": 1.0, "quantize": True, "aux": []}

    model = Model()
    hook = FusedMcaqHook(model, 4)
    x = torch.from_numpy(feature_map("smooth", B, C, H, H, 41)).cuda()
    with torch.no_grad():
        out = hook(None, None, x)
    rec = model._mcaq_state["aux"][0]
    assert out is rec["features_q"] and rec["layer"] == 4
    assert int((rec["bit_map"].cpu().numpy() != g["bit_map_mlp"]).sum()) <= 1


# ------------------------------------------------------------------ non-monotone mappers
def test_nonmonotone_mapper_is_evaluated_not_tabulated(M, W):
    from mcaq_yolo_b200 import constants as K, fused
    g = r2("nonmono_c3")
    mw = {k[len("mapper."):]: g[k] for k in g.files if k.startswith("mapper.")}
    m = M.ComplexityToBitMappingNetwork(enforce_monotonicity=False).cuda()
    m.load_state_dict(sd(mw))
    m.eval()
    assert not K.mapping_is_monotone(m)
    ramp = torch.linspace(0, 1, 400).reshape(1, 20, 20).cuda()
    with torch.no_grad():
        b = m(ramp, 1.0)
    pre = g["ramp_bits_cont"]
    amb = bit_ambiguous(pre)
    assert int(((b.cpu().numpy() != g["ramp_bits"]) & ~amb).sum()) == 0
    d = np.diff(g["ramp_bits"].ravel())
    assert (d > 0).any() and (d < 0).any()
    # fused hook: no step table for this network
    a, _, q = M.build_fixture_modules(W, "cuda")
    x = torch.from_numpy(feature_map("smooth", 1, 64, 80, 80, 43)).cuda()
    with torch.no_grad():
        rec, _ = fused.fused_scale_forward(x, a, m, q, 1.0, None)
    amb = bit_ambiguous(g["bit_map_cont"])
    assert int(((rec["bit_map"].cpu().numpy() != g["bit_map"]) & ~amb).sum()) == 0
    assert fused.mapper_block(m, 1.0).numel() == K.MAPPER_FLOATS


def test_constrained_mapper_with_negative_weights_falls_back(M, W):
    """enforce_monotonicity=True but a checkpoint saved before enforce_weight_constraints(): the step
    table must not be used until the constraint actually holds."""
    from mcaq_yolo_b200 import constants as K, fused
    _, m, _ = M.build_fixture_modules(W, "cuda")
    assert K.mapping_is_monotone(m)
    assert fused.mapper_block(m, 1.0).numel() == K.MAPPER_FLOATS + K.MAPPER_STEPS_FLOATS
    with torch.no_grad():
        m.mapping_network[3].weight[0, 0] = -0.25
    assert not K.mapping_is_monotone(m)
    assert fused.mapper_block(m, 1.0).numel() == K.MAPPER_FLOATS
    m.enforce_weight_constraints()
    assert K.mapping_is_monotone(m)


# ------------------------------------------------------------------ fp16 (autocast) feature maps
@pytest.mark.parametrize("name", ["c3_v8n_smooth", "c5_v8n_smooth"])
def test_fp16_feature_maps(name, M, W):
    from mcaq_yolo_b200 import fused
    c = Case(name)
    xh = torch.from_numpy(c.x()).cuda().half()
    a, m, q = M.build_fixture_modules(W, "cuda", grid_size=c.grid)
    with torch.no_grad():
        rec, _ = fused.fused_scale_forward(xh, a, m, q, 1.0, None)
        rec2 = M.mcaq_hook_forward(xh, a, m, q, temperature=1.0)
    assert rec["features_q"].dtype == torch.float16
    assert torch.equal(rec["features_q"], rec2["features_q"]) and torch.equal(rec["bit_map"], rec2["bit_map"])
    r = o.hook_forward(xh.float().cpu().numpy(), W["analyzer"], W["mapper"], W["quantizer"], c.grid, 1.0)
    assert np.array_equal(rec["bit_map"].cpu().numpy(), r["bit_map"])
    yo = torch.from_numpy(r["y"]).half()
    same = (rec["features_q"].cpu() == yo).float().mean().item()
    assert same > 0.9995, f"only {same:.5f} of y equals the fp16-rounded oracle value"
    np.testing.assert_allclose(rec["features_q"].float().cpu().numpy(), r["y"], rtol=2e-3, atol=1e-3)


def test_fp16_training_kernels_and_autocast_hook(M, W):
    """Under autocast the hooked outputs are fp16 (train.py:582): forward, backward and the module path."""
    from mcaq_yolo_b200 import ops
    c = Case("small_smooth")
    xs, g_, bf = c.x(), c.grad(), c["bit_map_frac"]
    mn, mx, m = c["train_run_min"], c["train_run_max"], c["train_soft_mask"]
    xh, gh = dev(xs).half(), dev(g_).half()
    qt = ops.build_qtable(None, dev(mn), dev(mx))
    y = ops.tile_quantize_train_fwd(xh, dev(bf), qt, dev(m))
    dx, dbit, dm = ops.tile_quantize_train_bwd(gh, xh, dev(bf), qt, dev(m))
    xu, gu = xh.float().cpu().numpy(), gh.float().cpu().numpy()
    yo = o.quantize_train_fwd(xu, bf, mn, mx, m)[0]
    dxo, dbo, dmo = o.quantize_train_bwd(gu, xu, bf, mn, mx, m)
    assert torch.equal(y.cpu(), torch.from_numpy(yo).half())
    assert torch.equal(dx.cpu(), torch.from_numpy(dxo).half())
    np.testing.assert_allclose(dbit.cpu().numpy(), dbo, rtol=2e-3, atol=2e-3 * np.abs(dbo).max())
    np.testing.assert_allclose(dm.cpu().numpy(), dmo, rtol=2e-3, atol=2e-3 * np.abs(dmo).max())
    a, mp, q = M.build_fixture_modules(W, "cuda")
    conv = torch.nn.Conv2d(16, 16, 1).cuda()
    x32 = dev(xs)
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.float16):
        feat = conv(x32)
        assert feat.dtype == torch.float16
        rec = M.mcaq_hook_forward(feat, a, mp, q, temperature=1.0)
    assert rec["features_q"].dtype == torch.float16 and torch.isfinite(rec["features_q"].float()).all()


# ------------------------------------------------------------------ the ambiguity set is empty
@pytest.mark.parametrize("name", CASE_NAMES)
def test_no_ambiguous_tiles_in_committed_cases(name, M, W):
    """The parity tests allow bit maps to differ inside the ambiguity set (continuous bit value within 1e-4
    of a .5 boundary).  On every committed case that set is EMPTY and the bit maps are simply equal."""
    c = Case(name)
    x = c.x()
    d = {}
    o.hook_forward(x, W["analyzer"], W["mapper"], W["quantizer"], c.grid, 1.0, detail=d)
    assert int(bit_ambiguous(d["bits_pre_round"]).sum()) == 0
    a, m, q = M.build_fixture_modules(W, "cuda", grid_size=c.grid)
    with torch.no_grad():
        rec = M.mcaq_hook_forward(torch.from_numpy(x).cuda(), a, m, q, temperature=1.0)
    assert np.array_equal(rec["bit_map"].cpu().numpy(), c["bit_map_mlp"])


# ------------------------------------------------------------------ Level-0 launcher
def test_level0_launcher_vector_path_and_error_channel(W):
    """`launch_spatial_quantization` (ops/src/mcaq_kernel.cu:102-111): routed to the vector kernel, equal to the
    table path; inconsistent tile arguments surface as an error (the reference's void launcher had none)."""
    from mcaq_yolo_b200 import _lib, ops
    c = Case("c3_v8n_smooth")
    x = dev(c.x())
    bm = dev(c["bit_map_rand"])
    mn, mx = dev(c["ch_min"]), dev(c["ch_max"])
    m = dev(c["soft_mask"]).unsqueeze(1)
    y = ops.spatial_quantize(x, bm, mn.view(1, -1, 1, 1), mx.view(1, -1, 1, 1), c.tile, c.tile, m)
    yo, _ = o.quantize_eval(c.x(), c["bit_map_rand"], c["ch_min"], c["ch_max"], c["soft_mask"])
    assert np.array_equal(y.cpu().numpy(), yo)
    assert _lib.load().mcaq_level0_status() == 0
    with pytest.raises(RuntimeError, match="launch_spatial_quantization"):
        ops.spatial_quantize(x, bm, mn.view(1, -1, 1, 1), mx.view(1, -1, 1, 1), c.tile + 1, c.tile, m)
    with pytest.raises(RuntimeError):
        ops.spatial_quantize(x, bm, mn[:3], mx[:3], c.tile, c.tile, m)


# ------------------------------------------------------------------ K3 with bulk-copy (TMA) staged input
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16, torch.float16])
@pytest.mark.parametrize("shape", [(3, 64, 80, 80), (2, 128, 40, 40), (5, 256, 20, 20), (1, 128, 160, 160)])
def test_k3_tma_variant_is_bit_identical(shape, dtype):
    """tile_quantize_tma_kernel (cp.async.bulk + mbarrier staging) against tile_quantize_vec_kernel: same bytes."""
    from mcaq_yolo_b200 import ops
    B, C, H, Wd = shape
    torch.manual_seed(11)
    x = (torch.randn(B, C, H, Wd, device="cuda") * 2 + 0.3).to(dtype)
    tile = ops.tile_size(H, 8)
    ht = H // tile
    bm = torch.randint(2, 9, (B, ht, ht), device="cuda").float()
    mask = torch.rand(B, H, Wd, device="cuda") * 0.3 + 0.7
    _, _, keys = ops.reduce_planes(x)
    packed = ops.ranges_decode(keys)
    saved = ops.K3_TMA
    try:
        for mk in (mask, None):
            ops.K3_TMA = False
            y0 = ops.tile_quantize_ranges(x, bm, packed, None, None, mk)
            ops.K3_TMA = True
            y1 = ops.tile_quantize_ranges(x, bm, packed, None, None, mk)
            torch.cuda.synchronize()
            assert torch.equal(y0, y1)
    finally:
        ops.K3_TMA = saved
