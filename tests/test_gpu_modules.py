"""GPU tests of the host-side mirror (mcaq_yolo_b200.modules): the reference's module/hook API
running on the sm_100a kernels, against the oracle and the committed reference outputs."""
import numpy as np
import pytest
import torch

import mcaq_oracle as o
from golden_util import Case, bit_ambiguous, sha, weights

pytestmark = pytest.mark.gpu
RTOL, ATOL = 1e-4, 2e-6

CASES = ["c3_v8n_smooth", "c4_v8n_smooth", "c5_v8n_smooth", "crop_50", "rect_24x40", "c3_grid4", "c3_grid16"]


@pytest.fixture(scope="module")
def M():
    from mcaq_yolo_b200 import modules
    return modules


@pytest.fixture(scope="module")
def W():
    return weights()


@pytest.mark.parametrize("name", CASES)
def test_hook_forward_matches_reference(name, M, W):
    c = Case(name)
    x = c.x()
    a, m, q = M.build_fixture_modules(W, "cuda", grid_size=c.grid)
    with torch.no_grad():
        rec = M.mcaq_hook_forward(torch.from_numpy(x).cuda(), a, m, q, temperature=1.0, layer=4)
    d = {}
    r = o.hook_forward(x, W["analyzer"], W["mapper"], W["quantizer"], c.grid, 1.0, detail=d)
    amb = bit_ambiguous(d["bits_pre_round"])
    bm = rec["bit_map"].cpu().numpy()
    assert int(((bm != c["bit_map_mlp"]) & ~amb).sum()) == 0
    np.testing.assert_allclose(rec["complexity"].cpu().numpy(), c["complexity"], rtol=RTOL, atol=ATOL)
    y = rec["features_q"].cpu().numpy()
    if np.array_equal(bm, c["bit_map_mlp"]):
        np.testing.assert_allclose(y[:, ::3, ::5, ::7], c["y_sub"], rtol=RTOL, atol=ATOL)
        np.testing.assert_allclose(y, r["y"], rtol=RTOL, atol=ATOL)
    assert rec["layer"] == 4 and y.shape == x.shape


def test_linear_mapper_and_score_image(M, W):
    c = Case("c3_v8n_smooth")
    x = torch.from_numpy(c.x()).cuda()
    a, m, q = M.build_fixture_modules(W, "cuda", linear_mapper=True)
    with torch.no_grad():
        cpx = a(x)
        b = m(cpx, 1.0)
        s = a.score_image(x)
    assert np.array_equal(b.cpu().numpy(), c["bit_map_linear"])
    np.testing.assert_allclose(s.cpu().numpy(), c["score"], rtol=RTOL, atol=ATOL)
    # reference known answers (tests/test_smoke.py:199-211)
    flat = torch.full((1, 8, 8), 0.5, device="cuda")
    assert torch.all(m(flat) == 5)
    assert float(m(torch.zeros(1, 8, 8, device="cuda")).max()) == 2
    assert float(m(torch.ones(1, 8, 8, device="cuda")).min()) == 8


def test_frozen_calibration_and_ema(M, W):
    """calibrate (training=True under eval, models/mcaq_yolo.py:446) -> freeze -> frozen ranges."""
    c = Case("c4_v8n_smooth")
    x = c.x()
    a, m, q = M.build_fixture_modules(W, "cuda")
    xt = torch.from_numpy(x).cuda()
    bm = torch.from_numpy(c["bit_map_rand"]).cuda()
    with torch.no_grad():
        q(xt, bm, training=True)
        assert np.array_equal(q.running_min.cpu().numpy().ravel(), c["train_run_min"])
        q(xt * 1.5 + 0.25, bm, training=True)
    np.testing.assert_allclose(q.running_min.cpu().numpy().ravel(), c["ema2_min"], rtol=1e-6)
    assert int(q.num_batches_tracked) == 2
    q.freeze_calibration()
    frozen = q.running_min.clone()
    with torch.no_grad():
        q(xt * 100, bm, training=True)
        assert torch.equal(q.running_min, frozen), "stats moved after freeze"
        y = q(xt, bm, training=False)
    mn, mx = q.running_min.cpu().numpy().ravel(), q.running_max.cpu().numpy().ravel()
    _, ab = o.channel_sums(x)
    mo = o.soft_mask(c["bit_map_rand"], ab, c.C, W["quantizer"])
    yo, _ = o.quantize_eval(x, c["bit_map_rand"], mn, mx, mo)
    np.testing.assert_allclose(y.cpu().numpy(), yo, rtol=RTOL, atol=ATOL)
    # checkpoint round trip keeps the lazily created buffers (quantization.py:297-312)
    q2 = M.SpatialAdaptiveQuantization().cuda()
    q2.load_state_dict(q.state_dict())
    assert torch.equal(q2.running_min, q.running_min) and q2._is_frozen()


def test_training_step_gradients(M, W):
    """Train-mode hook: continuous bits, fractional quantiser, gradients reach x, the bit map
    (fractional + soft-mask paths) and the soft-mask net; compared with the reference's autograd."""
    c = Case("small_smooth")
    x = torch.from_numpy(c.x()).cuda().requires_grad_(True)
    g = torch.from_numpy(c.grad()).cuda()
    a, m, q = M.build_fixture_modules(W, "cuda")
    q.train()
    bf = torch.from_numpy(c["bit_map_frac"]).cuda().requires_grad_(True)
    y = q(x, bf, training=True)
    y.backward(g)
    np.testing.assert_allclose(y.detach().cpu().numpy()[:, ::3, ::5, ::7], c["train_y_sub"], rtol=RTOL, atol=ATOL)
    np.testing.assert_allclose(x.grad.cpu().numpy()[:, ::3, ::5, ::7], c["train_dx_sub"], rtol=RTOL, atol=ATOL)
    ref = c["train_dbit_total"]
    np.testing.assert_allclose(bf.grad.cpu().numpy(), ref, rtol=5e-3, atol=5e-3 * np.abs(ref).max())
    assert q.soft_mask.net[0].weight.grad is not None and float(q.soft_mask.net[0].weight.grad.abs().sum()) > 0
    # analyzer + mapper in train mode: gradient reaches both MLPs (tests/test_smoke.py:50-59, 87-96)
    a.train(); m.train()
    cpx = a(x.detach())
    bits = m(cpx, 10.0, return_continuous=True)
    (bits.mean() - 4.0).pow(2).backward()
    assert any(p.grad is not None and float(p.grad.abs().sum()) > 0 for p in a.complexity_mlp.parameters())
    assert any(p.grad is not None and float(p.grad.abs().sum()) > 0 for p in m.mapping_network.parameters())


def test_cuda_graph_capture_and_bf16(M, W):
    """The whole eval hook is capturable (no host syncs) and bf16 maps keep codes exact."""
    c = Case("c3_v8n_smooth")
    xb = torch.from_numpy(c.x()).cuda().to(torch.bfloat16)
    a, m, q = M.build_fixture_modules(W, "cuda")
    with torch.no_grad():
        eager = M.mcaq_hook_forward(xb, a, m, q)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            rec = M.mcaq_hook_forward(xb, a, m, q)
        g.replay()
        torch.cuda.synchronize()
    assert torch.equal(rec["bit_map"], eager["bit_map"]) and torch.equal(rec["features_q"], eager["features_q"])
    xu = xb.float().cpu().numpy()
    r = o.hook_forward(xu, W["analyzer"], W["mapper"], W["quantizer"], 8, 1.0)
    assert np.array_equal(rec["bit_map"].cpu().numpy(), r["bit_map"])
    # y is the bf16 rounding of the fp32 value the oracle computes on the upcast input: equal, bit for bit
    assert torch.equal(rec["features_q"].cpu(), torch.from_numpy(r["y"]).to(torch.bfloat16))


def test_reference_cuda_parity_test_shape(M):
    """The reference's own parity test (tests/test_smoke.py:226-246) through the shimmed
    `mcaq_cuda_ops.spatial_quantize`: randn(2,8,32,32), randint bit map, atol 1e-4."""
    mod = M.install_mcaq_cuda_ops()
    import mcaq_cuda_ops
    assert mcaq_cuda_ops is mod
    torch.manual_seed(0)
    for smooth in (False, True):
        q = M.SpatialAdaptiveQuantization(smooth_transitions=smooth).cuda().eval()
        x = torch.randn(2, 8, 32, 32, device="cuda")
        bit_map = torch.randint(2, 9, (2, 4, 4), device="cuda").float()
        with torch.no_grad():
            mask = q.soft_mask(bit_map, x).float().contiguous() if smooth else None
            out = mcaq_cuda_ops.spatial_quantize(x.contiguous(), bit_map, x.amin(dim=(0, 2, 3), keepdim=True),
                                                 x.amax(dim=(0, 2, 3), keepdim=True), 8, 8, mask)
            mn = x.amin(dim=(0, 2, 3)).cpu().numpy()
            mx = x.amax(dim=(0, 2, 3)).cpu().numpy()
            yo, _ = o.quantize_eval(x.cpu().numpy(), bit_map.cpu().numpy(), mn, mx,
                                    None if mask is None else mask[:, 0].cpu().numpy())
        assert np.allclose(out.cpu().numpy(), yo, atol=1e-4)
        assert np.array_equal(out.cpu().numpy(), yo)


def test_install_swaps_modules(M, W):
    class Fake(torch.nn.Module):
        def __init__(self):
            super().__init__()
            a, m, q = M.build_fixture_modules(W, "cuda")
            self.complexity_analyzer, self.bit_mapper = a, m
            self.quantizers = torch.nn.ModuleDict({"4": q})
    f = Fake()
    before = {k: v.clone() for k, v in f.state_dict().items()}
    M.install(f)
    after = f.state_dict()
    assert set(before) == set(after) and all(torch.equal(before[k], after[k]) for k in before)


# ------------------------------------------------------------------------------ fused 3-launch path
@pytest.mark.parametrize("name", CASES + ["c3_v8s1280_smooth", "small_noise"])
@pytest.mark.parametrize("linear", [False, True])
def test_fused_hook_equals_module_path(name, linear, M, W):
    """K1 -> K2 fused -> K3 (three launches) must reproduce the module-by-module path bit for bit,
    including the range-key re-arming across consecutive calls."""
    from mcaq_yolo_b200 import fused
    c = Case(name)
    a, m, q = M.build_fixture_modules(W, "cuda", grid_size=c.grid, linear_mapper=linear)
    x = torch.from_numpy(c.x()).cuda()
    x2 = (x * 0.5 - 0.25).contiguous()
    with torch.no_grad():
        ref1 = M.mcaq_hook_forward(x, a, m, q, temperature=1.0)
        ref2 = M.mcaq_hook_forward(x2, a, m, q, temperature=1.3)
        rec1, ws = fused.fused_scale_forward(x, a, m, q, 1.0, None)
        rec2, ws = fused.fused_scale_forward(x2, a, m, q, 1.3, ws)     # re-armed keys
    for ref, rec in ((ref1, rec1), (ref2, rec2)):
        assert torch.equal(ref["complexity"], rec["complexity"])
        assert torch.equal(ref["bit_map"], rec["bit_map"])
        assert torch.equal(ref["features_q"], rec["features_q"])
    if not linear:
        assert int((rec1["bit_map"].cpu().numpy() != c["bit_map_mlp"]).sum()) == 0


def test_fused_frozen_and_no_mask(M, W):
    from mcaq_yolo_b200 import fused
    c = Case("c4_v8n_smooth")
    x = torch.from_numpy(c.x()).cuda()
    a, m, q = M.build_fixture_modules(W, "cuda")
    with torch.no_grad():
        q(x, torch.from_numpy(c["bit_map_rand"]).cuda(), training=True)     # one calibration batch
        q.freeze_calibration()
        ref = M.mcaq_hook_forward(x * 1.1, a, m, q)
        rec, _ = fused.fused_scale_forward(x * 1.1, a, m, q, 1.0, None)
        assert torch.equal(ref["features_q"], rec["features_q"])
        q2 = M.SpatialAdaptiveQuantization(smooth_transitions=False).cuda().eval()
        ref = M.mcaq_hook_forward(x, a, m, q2)
        rec, _ = fused.fused_scale_forward(x, a, m, q2, 1.0, None)
        assert torch.equal(ref["features_q"], rec["features_q"]) and torch.equal(ref["bit_map"], rec["bit_map"])


def test_fused_hot_path_streams_and_graph(M, W):
    from mcaq_yolo_b200 import fused
    names = ["c3_v8n_smooth", "c4_v8n_smooth", "c5_v8n_smooth"]
    feats = [torch.from_numpy(Case(n).x()).cuda().to(torch.bfloat16) for n in names]
    a, m, _ = M.build_fixture_modules(W, "cuda")
    qs = [M.build_fixture_modules(W, "cuda")[2] for _ in names]
    serial = fused.FusedHotPath(a, m, qs, streams=False)
    multi = fused.FusedHotPath(a, m, qs, streams=True)
    r0 = serial.run(feats)
    r1 = multi.run(feats)
    torch.cuda.synchronize()
    for u, v in zip(r0, r1):
        assert torch.equal(u["features_q"], v["features_q"]) and torch.equal(u["bit_map"], v["bit_map"])
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        r2 = multi.run(feats)
    g.replay()
    g.replay()
    torch.cuda.synchronize()
    for u, v in zip(r0, r2):
        assert torch.equal(u["features_q"], v["features_q"]) and torch.equal(u["bit_map"], v["bit_map"])


def test_install_registers_fused_hooks(M, W):
    """A model-shaped object with the reference's hook protocol: install() swaps the hooks and a
    forward produces the reference's aux records through the three-launch path."""
    from mcaq_yolo_b200 import fused

    class Backbone(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.model = torch.nn.Sequential(torch.nn.Identity(), torch.nn.Identity())

        def forward(self, x):
            return self.model(x)

    class FakeMcaq(torch.nn.Module):
        def __init__(self):
            super().__init__()
            a, m, q = M.build_fixture_modules(W, "cuda")
            self.complexity_analyzer, self.bit_mapper = a, m
            self.quantizers = torch.nn.ModuleDict({"1": q})
            self.model = Backbone()
            self.backbone_out_indices = [1]
            self._mcaq_hooks = [self.model.model[1].register_forward_hook(lambda mod, i, o: None)]
            self._mcaq_state = {"active": False}
            self.normalize_complexity = False

    f = FakeMcaq().eval()
    M.install(f)
    assert isinstance(list(f.model.model[1]._forward_hooks.values())[0], fused.FusedMcaqHook)
    c = Case("c3_v8n_smooth")
    x = torch.from_numpy(c.x()).cuda()
    f._mcaq_state = {"active": True, "temperature": 1.0, "quantize": True, "aux": []}
    with torch.no_grad():
        y = f.model(x)
    aux = f._mcaq_state["aux"]
    assert len(aux) == 1 and aux[0]["layer"] == 1 and torch.equal(y, aux[0]["features_q"])
    assert np.array_equal(aux[0]["bit_map"].cpu().numpy(), c["bit_map_mlp"])
    np.testing.assert_allclose(y.cpu().numpy()[:, ::3, ::5, ::7], c["y_sub"], rtol=RTOL, atol=ATOL)


def test_sharded_hot_path_single_rank_equals_fused(M, W):
    """The multi-rank phase split (K1+decode | all-reduce | K2 | K3) with world_size 1 reproduces the
    fused path bit for bit, eagerly and as three captured graphs."""
    from mcaq_yolo_b200 import fused
    names = ["c3_v8n_smooth", "c4_v8n_smooth", "c5_v8n_smooth"]
    cases = [Case(n) for n in names]
    feats = [torch.from_numpy(c.x()).cuda() for c in cases]
    a, m, _ = M.build_fixture_modules(W, "cuda")
    qs = [M.build_fixture_modules(W, "cuda")[2] for _ in names]
    ref = fused.FusedHotPath(a, m, qs, streams=False).run(feats)
    sh = fused.ShardedHotPath(a, m, qs, [(c.C, c.H, c.W) for c in cases], torch.device("cuda"))
    out = sh.run(feats)
    torch.cuda.synchronize()
    for u, v in zip(ref, out):
        assert torch.equal(u["features_q"], v["features_q"]) and torch.equal(u["bit_map"], v["bit_map"])
    gA, gB1, gB2 = torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph()
    with torch.no_grad():
        with torch.cuda.graph(gA):
            planes = sh.sweep(feats)
        with torch.cuda.graph(gB1):
            nets_out = sh.nets(feats, planes)
        with torch.cuda.graph(gB2):
            recs = sh.quantize(feats, nets_out)
    for _ in range(2):
        gA.replay()
        done = sh.exchange()
        gB1.replay()
        torch.cuda.current_stream().wait_event(done)
        gB2.replay()
    torch.cuda.synchronize()
    for u, v in zip(ref, recs):
        assert torch.equal(u["features_q"], v["features_q"]) and torch.equal(u["bit_map"], v["bit_map"])


# ------------------------------------------------------------------------------ K2 cluster split
@pytest.mark.parametrize("name", CASES + ["c3_v8s1280_smooth", "small_noise"])
@pytest.mark.parametrize("linear", [False, True])
def test_cluster_split_is_bit_identical(name, linear, M, W):
    """The morphology kernel splits an image's tile rows over a 1/2/4-CTA cluster (DSMEM
    all-gathers); every split must give the same phi / complexity / bit map / mask bit for bit."""
    from mcaq_yolo_b200 import _lib, constants as K, ops
    lib = _lib.load()
    c = Case(name)
    a, m, q = M.build_fixture_modules(W, "cuda", grid_size=c.grid, linear_mapper=linear)
    cm = K.pack_complexity_mlp(a.complexity_mlp)
    mp = None if linear else K.pack_mapping_network(m.mapping_network)
    sm = K.pack_soft_mask(q.soft_mask)
    x = torch.from_numpy(c.x()).cuda()
    s, ab, _ = ops.reduce_planes(x)
    outs = []
    try:
        for ns in (0, 1, 2, 4, 8):
            lib.mcaq_debug_cluster_split(ns)
            try:
                r = ops.morph_fused(s, ab, c.C, c.grid, cm, mp, sm, 1.0, want_phi=True)
            except RuntimeError as e:           # a forced split may not fit shared memory (160x160 on one CTA)
                assert ns != 0 and "on-chip budget" in str(e)
                continue
            torch.cuda.synchronize()
            outs.append(r)
    finally:
        lib.mcaq_debug_cluster_split(0)
    for r in outs[1:]:
        for k in ("phi", "complexity", "bit_map", "mask"):
            assert torch.equal(outs[0][k], r[k]), f"{k} differs between cluster splits"
    if not linear:
        assert int((outs[0]["bit_map"].cpu().numpy() != c["bit_map_mlp"]).sum()) == 0


# ------------------------------------------------------------------------------ multi-GPU range exchange
def test_peer_range_exchange_virtual_ranks(M, W):
    """The in-kernel range exchange (csrc/peer_exchange.cuh) driven by two virtual ranks on one GPU
    (one stream per rank: each rank's morphology kernel publishes at its start and waits for the other
    at its end): a batch split in two must quantize exactly like the unsharded batch, over several
    steps (slot / flag double buffering); the host-driven publish / merge halves must agree with
    torch.minimum."""
    from mcaq_yolo_b200 import fused
    from mcaq_yolo_b200.peer import RangeExchange
    c = Case("c4_v8n_smooth")
    a, m, q = M.build_fixture_modules(W, "cuda", grid_size=c.grid)
    x = torch.from_numpy(c.x()).cuda()
    x = torch.cat([x, x * 0.7 + 0.2, x * 1.3 - 0.1, -x], 0).contiguous()        # different ranges per shard
    half = x.shape[0] // 2
    shards = [x[:half].contiguous(), x[half:].contiguous()]
    ex = RangeExchange.virtual(c.C, 2)
    ws = [None, None]
    streams = [torch.cuda.Stream(), torch.cuda.Stream()]
    for step in range(3):
        xs = [s * (1.0 + 0.25 * step) for s in shards]
        full = torch.cat(xs, 0)
        with torch.no_grad():
            ref, _ = fused.fused_scale_forward(full, a, m, q, 1.0, None)
        torch.cuda.synchronize()
        recs = [None, None]
        for r in range(2):
            with torch.cuda.stream(streams[r]), torch.no_grad():
                recs[r], ws[r] = fused.fused_scale_forward(xs[r], a, m, q, 1.0, ws[r], xchg=ex[r])
        torch.cuda.synchronize()
        assert torch.equal(torch.cat([r_["features_q"] for r_ in recs], 0), ref["features_q"]), \
            f"step {step}: sharded != unsharded"
        assert torch.equal(torch.cat([r_["bit_map"] for r_ in recs], 0), ref["bit_map"])
    # host-driven halves
    ex3 = RangeExchange.virtual(8, 3)
    vecs = [torch.randn(16, device="cuda") for _ in range(3)]
    for e, v in zip(ex3, vecs):
        e.publish(v)
    want = torch.minimum(torch.minimum(vecs[0], vecs[1]), vecs[2])
    for e in ex3:
        assert torch.equal(e.merged("cuda"), want)


# ------------------------------------------------------------------------------ channels_last (NHWC) inputs
@pytest.mark.parametrize("name", ["c3_v8n_smooth", "c4_v8n_smooth", "c5_v8n_smooth", "c3_v8s1280_smooth"])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_channels_last_matches_nchw(name, dtype, M, W):
    """channels_last feature maps run through native NHWC forms of K1 and K3 (no layout copy); planes,
    ranges, bit maps and the quantized map must equal the NCHW path bit for bit, and the output keeps
    the input's memory format."""
    from mcaq_yolo_b200 import fused, ops
    c = Case(name)
    a, m, q = M.build_fixture_modules(W, "cuda", grid_size=c.grid)
    x = torch.from_numpy(c.x()).cuda().to(dtype)
    xcl = x.contiguous(memory_format=torch.channels_last)
    assert ops.is_nhwc(xcl) and not ops.is_nhwc(x)
    s0, a0, k0 = ops.reduce_planes(x)
    s1, a1, k1 = ops.reduce_planes(xcl)
    assert torch.equal(s0, s1) and torch.equal(a0, a1) and torch.equal(k0, k1)
    with torch.no_grad():
        ref, _ = fused.fused_scale_forward(x, a, m, q, 1.0, None)
        rec, _ = fused.fused_scale_forward(xcl, a, m, q, 1.0, None)
    assert rec["features_q"].is_contiguous(memory_format=torch.channels_last)
    assert torch.equal(ref["bit_map"], rec["bit_map"]) and torch.equal(ref["complexity"], rec["complexity"])
    assert torch.equal(ref["features_q"], rec["features_q"])
    # frozen calibration and no soft mask
    with torch.no_grad():
        q(x.float(), ref["bit_map"], training=True)
        q.freeze_calibration()
        ref2, _ = fused.fused_scale_forward(x, a, m, q, 1.0, None)
        rec2, _ = fused.fused_scale_forward(xcl, a, m, q, 1.0, None)
    assert torch.equal(ref2["features_q"], rec2["features_q"])
