"""CPU tests of the host side: interface parity with the reference modules (state_dict keys,
constructor signatures), the generated constants header, batch sharding and the world-size-2
range merge over gloo.  No kernel is launched here."""
import inspect
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))

from mcaq_yolo_b200 import constants as K
from mcaq_yolo_b200 import modules as M
import mcaq_oracle as o


def _reference():
    from ref_loader import load_reference, reference_root
    if reference_root() is None:
        pytest.skip("reference tree not mounted (GPU box): interface parity is pinned in the build container")
    return load_reference()


def test_state_dict_keys_match_reference():
    """reference checkpoints must load -- same keys and shapes (incl. lazily created
    running_min/max, quantization.py:297-312)."""
    morph, ba, qz = _reference()
    pairs = [
        (morph.MorphologicalComplexityAnalyzer(device="cpu"), M.MorphologicalComplexityAnalyzer(device="cpu")),
        (ba.ComplexityToBitMappingNetwork(), M.ComplexityToBitMappingNetwork()),
        (qz.SpatialAdaptiveQuantization(), M.SpatialAdaptiveQuantization()),
        (qz.LearnedSoftMask(), M.LearnedSoftMask()),
    ]
    for ref, mine in pairs:
        rs, ms = ref.state_dict(), mine.state_dict()
        assert list(rs.keys()) == list(ms.keys()), type(ref).__name__
        for k in rs:
            assert rs[k].shape == ms[k].shape and rs[k].dtype == ms[k].dtype, k
        mine.load_state_dict(rs, strict=True)
    # a calibrated reference quantizer (running stats present) loads strictly into a fresh native one
    q = qz.SpatialAdaptiveQuantization()
    q.train()
    q(torch.randn(2, 4, 16, 16), torch.full((2, 4, 4), 4.0), training=True)
    q.freeze_calibration()
    mine = M.SpatialAdaptiveQuantization()
    mine.load_state_dict(q.state_dict(), strict=True)
    assert torch.equal(mine.running_min, q.running_min) and mine._is_frozen()
    # soft-mask smoothing buffer and init are identical
    assert torch.equal(qz.LearnedSoftMask().smooth_kernel, M.LearnedSoftMask().smooth_kernel)


def test_constructor_and_forward_signatures_match_reference():
    morph, ba, qz = _reference()
    for ref, mine in ((morph.MorphologicalComplexityAnalyzer, M.MorphologicalComplexityAnalyzer),
                      (ba.ComplexityToBitMappingNetwork, M.ComplexityToBitMappingNetwork),
                      (ba.LinearBitMapper, M.LinearBitMapper),
                      (qz.LearnedSoftMask, M.LearnedSoftMask)):
        assert list(inspect.signature(ref.__init__).parameters) == list(inspect.signature(mine.__init__).parameters)
        assert list(inspect.signature(ref.forward).parameters) == list(inspect.signature(mine.forward).parameters)
    ref_q = list(inspect.signature(qz.SpatialAdaptiveQuantization.__init__).parameters)
    my_q = list(inspect.signature(M.SpatialAdaptiveQuantization.__init__).parameters)
    assert my_q[:len(ref_q)] == ref_q          # extra trailing kwargs: process_group, sync_ranges
    for name in ("freeze_calibration", "update_running_stats", "enforce_weight_constraints", "score_image",
                 "compute_phi_tiles", "_tile_size", "bilateral_filter", "create_augmented_features"):
        assert any(hasattr(c, name) for c in (M.MorphologicalComplexityAnalyzer, M.ComplexityToBitMappingNetwork,
                                              M.SpatialAdaptiveQuantization))


def test_unsupported_variants_fail_loudly():
    with pytest.raises(NotImplementedError):
        M.MorphologicalComplexityAnalyzer(device="cpu", metric_backend="cv2")
    with pytest.raises(NotImplementedError):
        M.SpatialAdaptiveQuantization(calibration_mode="percentile")
    a = M.MorphologicalComplexityAnalyzer(device="cpu")
    with pytest.raises(RuntimeError):
        a(torch.rand(1, 3, 32, 32))              # CPU tensor: no fallback


def test_tile_size_rule():
    a = M.MorphologicalComplexityAnalyzer(device="cpu", grid_size=8)
    for H in (640, 160, 80, 50, 40, 20, 7):
        assert a._tile_size(H) == o.tile_size(H, 8)
    assert [M.MorphologicalComplexityAnalyzer(device="cpu", grid_size=g)._tile_size(160) for g in (4, 8, 16)] == [32, 16, 8]


def test_constant_block_and_packers_match_oracle():
    c = K.constant_block()
    assert np.array_equal(c[0:25], o.CANNY_BLUR.ravel()) and np.array_equal(c[25:146], o.ADAPT_BLUR.ravel())
    assert np.array_equal(c[146:171], o.BILATERAL_SPATIAL.ravel())
    _, x, w = o.fractal_tables(32)
    assert np.array_equal(c[171:176], x) and np.array_equal(c[176:181], w)
    assert c[181] == o.RAD2DEG and c[182] == o.FOUR_PI and c[183] == o.LOG2_10 and c[184] == o.BILATERAL_RANGE_DEN
    m = M.ComplexityToBitMappingNetwork()
    m.train()
    m(torch.rand(4, 6, 6))                       # non-trivial BN running stats
    m.eval()
    packed = K.pack_mapping_network(m.mapping_network).numpy()
    sd = {k: v.numpy() for k, v in m.state_dict().items()}
    # BN fold equals the oracle's batch_norm_eval affine
    x = np.linspace(-1, 1, 32, dtype=np.float32)[None, :]
    ref = o.batch_norm_eval(x, sd["mapping_network.1.weight"], sd["mapping_network.1.bias"],
                            sd["mapping_network.1.running_mean"], sd["mapping_network.1.running_var"])
    alpha, beta = packed[128:160], packed[160:192]
    assert np.array_equal((x * alpha).astype(np.float32) + beta, ref)
    assert packed.size == K.MAPPER_FLOATS
    assert K.pack_mapping_network(m.mapping_network) is K.pack_mapping_network(m.mapping_network)   # cached
    with torch.no_grad():
        m.mapping_network[0].weight.add_(1.0)
    assert not np.array_equal(K.pack_mapping_network(m.mapping_network).numpy(), packed)             # invalidated


def test_generated_consts_header_is_current():
    import gen_consts_header
    path = os.path.join(ROOT, "mcaq_yolo_b200", "csrc", "mcaq_consts.cuh")
    assert open(path).read() == gen_consts_header.generate(), "run tools/gen_consts_header.py"


def test_shard_batch_partitions():
    for n, world in ((256, 8), (64, 3), (5, 8), (128, 2)):
        spans = [M.shard_batch(n, r, world) for r in range(world)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
        assert max(e - s for s, e in spans) - min(e - s for s, e in spans) <= 1


def _gloo_worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        sys.path.insert(0, os.path.join(ROOT, "oracle"))
        from inputs import feature_map
        x = feature_map("smooth", 6, 16, 24, 24, seed=3)
        s, e = M.shard_batch(x.shape[0], rank, world)
        mn, mx = o.channel_minmax(x[s:e])                     # this rank's shard
        packed = torch.from_numpy(np.concatenate([mn, -mx]).astype(np.float32))
        M.allreduce_ranges(packed)                            # the path's only collective
        gmn, gmx = o.channel_minmax(x)
        ok = np.array_equal(packed[:16].numpy(), gmn) and np.array_equal(-packed[16:].numpy(), gmx)
        # per-image quantities need no exchange: a shard's phi equals the same rows of the full batch
        phi_full = o.phi_tiles(x, 8)
        ok = ok and np.array_equal(o.phi_tiles(x[s:e], 8), phi_full[s:e])
        # with merged ranges the shard's codes equal the unsharded batch's codes
        bm = np.full((x.shape[0], 6, 6), 5.0, np.float32)
        _, codes_full = o.quantize_eval(x, bm, gmn, gmx, None)
        _, codes_shard = o.quantize_eval(x[s:e], bm[s:e], packed[:16].numpy(), -packed[16:].numpy(), None)
        ok = ok and np.array_equal(codes_shard, codes_full[s:e])
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


def test_range_merge_world_size_2_gloo():
    """Batch sharded over 2 ranks: one MIN all-reduce of [min, -max] reproduces the unsharded
    per-channel ranges, hence bit-identical codes (quantization.py:650-654 semantics kept)."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    assert res == [(0, True), (1, True)]


def test_kd_geometry_rule_matches_vector_path_contract():
    """ops.kd_geometry_ok mirrors the C-side test of the training vector path (include/mcaq_b200.h:
    H*W % VEC == 0, W % 4 == 0, W % Wt == 0, (W / Wt) % 4 == 0)."""
    import torch
    from mcaq_yolo_b200 import ops

    def ok(shape, wt, dtype=torch.float32):
        return ops.kd_geometry_ok(torch.empty(shape, dtype=dtype), torch.empty(shape[0], wt, wt))

    assert ok((2, 64, 80, 80), 10) and ok((2, 128, 40, 40), 10) and ok((2, 256, 20, 20), 5)
    assert ok((2, 256, 20, 20), 5, torch.bfloat16) and ok((1, 128, 160, 160), 10, torch.bfloat16)
    assert not ok((1, 3, 7, 9), 2) and not ok((2, 8, 50, 50), 12)
    assert not ok((1, 8, 12, 12), 6)                    # tile width 2: segments would straddle tiles
    assert ok((1, 8, 5, 12), 3) and not ok((1, 8, 5, 12), 3, torch.bfloat16)     # 60 pixels: multiple of 4, not of 8


def test_mlp_mapper_is_a_staircase_in_the_oracle():
    """The step table of the fused kernel (mcaq_mapper_steps) relies on the monotonic mapper of Eq.18
    giving a non-decreasing integer bit width in the scalar complexity: check it on the oracle's
    restatement of the reference mapper with the fixture weights, over a dense sorted sample."""
    import mcaq_oracle as o
    from golden_util import weights
    W = weights()
    rng = np.random.default_rng(11)
    c = np.sort(np.concatenate([rng.random(20000, dtype=np.float32), np.linspace(0, 1, 2049, dtype=np.float32)]))
    for t in (1.0, 0.7, 1.3):
        bits = o.mlp_bit_mapper(c.reshape(1, 1, -1), W["mapper"], temperature=t, continuous=False).reshape(-1)
        assert np.all(np.diff(bits) >= 0), "mapper output must be non-decreasing in complexity"
        assert set(np.unique(bits)) <= set(np.arange(2, 9, dtype=np.float32))
    assert np.unique(o.mlp_bit_mapper(c.reshape(1, 1, -1), W["mapper"], temperature=1.0, continuous=False)).size >= 4


# ---- training-time exchanges of the sharded batch over gloo, world size 2 -------------------------
def _gloo_train_worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from mcaq_yolo_b200 import train_nets as TN
        torch.manual_seed(5)
        ok = True
        # (1) SyncBN merge: rank-ordered Chan merge of per-rank (count, mean, M2) == statistics of the whole batch
        z = torch.randn(10, 16) * 3 + 1.5
        parts_idx = [(0, 7), (7, 10)]                        # uneven shards
        s, e = parts_idx[rank]
        mine = z[s:e]
        local = torch.cat([torch.tensor([float(e - s)]), mine.mean(0), ((mine - mine.mean(0)) ** 2).sum(0)])
        gathered = [torch.zeros_like(local) for _ in range(world)]
        dist.all_gather(gathered, local)
        cnt, mean, m2 = TN.merge_batch_stats([(float(g[0]), g[1:17], g[17:]) for g in gathered])
        ok = ok and cnt == 10 and torch.allclose(mean, z.mean(0), atol=1e-6) and \
            torch.allclose(m2 / cnt, z.var(0, unbiased=False), atol=1e-5)
        # (2) avg_bits / Lbit / Lsmooth over the GLOBAL batch with a local straight-through gradient
        maps_full = [torch.rand(4, 6, 6) * 6 + 2, torch.rand(4, 3, 3) * 6 + 2]

        class _CpuBitStats(torch.autograd.Function):          # CPU stand-in for the bit_stats kernel (same reductions)
            @staticmethod
            def forward(ctx, b):
                ctx.save_for_backward(b)
                dx = (b[:, 1:, :] - b[:, :-1, :]).abs().sum()
                dy = (b[:, :, 1:] - b[:, :, :-1]).abs().sum()
                return torch.stack([b.sum(), dx + dy])

            @staticmethod
            def backward(ctx, g):
                b, = ctx.saved_tensors
                with torch.enable_grad():
                    bb = b.detach().requires_grad_(True)
                    tv = (bb[:, 1:, :] - bb[:, :-1, :]).abs().sum() + (bb[:, :, 1:] - bb[:, :, :-1]).abs().sum()
                    (g[0] * bb.sum() + g[1] * tv).backward()
                return bb.grad

        TN.BitStatsFn = _CpuBitStats
        shard = [m[2 * rank:2 * rank + 2].clone().requires_grad_(True) for m in maps_full]
        avg, lbit, lsm = TN.bit_map_losses(shard, 4.0)
        (lbit + lsm).backward()
        full = [m.clone().requires_grad_(True) for m in maps_full]
        avg_f = torch.stack([m.mean() for m in full]).mean()
        tv = []
        for m in full:
            dx = (m[:, 1:, :] - m[:, :-1, :]).abs()
            dy = (m[:, :, 1:] - m[:, :, :-1]).abs()
            tv.append((dx.sum() + dy.sum()) / (dx.numel() + dy.numel()))
        ((avg_f - 4.0) ** 2 + sum(tv) / 2).backward()
        ok = ok and torch.allclose(avg, avg_f, atol=1e-6) and torch.allclose(lsm, sum(tv) / 2, atol=1e-6)
        for a, b in zip(shard, full):
            ok = ok and torch.allclose(a.grad, b.grad[2 * rank:2 * rank + 2], atol=1e-6)
        # (3) one flat all-reduce of the small networks' gradients
        lin = torch.nn.Linear(3, 2)
        for p in lin.parameters():
            p.grad = torch.full_like(p, float(rank + 1))
        TN.allreduce_grads([lin])
        ok = ok and all(torch.equal(p.grad, torch.full_like(p, 3.0)) for p in lin.parameters())
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


def test_training_exchanges_world_size_2_gloo():
    """SyncBN statistics merge, global avg_bits with local gradient, flat gradient all-reduce."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_gloo_train_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    assert res == [(0, True), (1, True)], res


def test_flat_params_is_shared_within_a_step_and_dropped_after_backward():
    """train_nets.flat_params(owner=...): one concatenation per training step (the three scales share the node, so
    autograd splits the summed flat gradient once), rebuilt after a backward pass or a parameter update; gradients
    equal the per-call concatenation's; the module stays deep-copyable."""
    import copy
    import torch
    from mcaq_yolo_b200 import train_nets as TN
    torch.manual_seed(0)
    m = torch.nn.Sequential(torch.nn.Linear(4, 3), torch.nn.Linear(3, 1))
    w = [torch.randn(19) for _ in range(3)]
    f = [TN.flat_params(m.parameters(), owner=m) for _ in range(3)]
    assert f[0] is f[1] and f[1] is f[2]
    sum((fi * wi).sum() for fi, wi in zip(f, w)).backward()
    got = [p.grad.clone() for p in m.parameters()]
    for p in m.parameters():
        p.grad = None
    sum((TN.flat_params(m.parameters()) * wi).sum() for wi in w).backward()        # no owner: a fresh cat per call
    for a, p in zip(got, m.parameters()):
        assert torch.allclose(a, p.grad, rtol=1e-6, atol=1e-6)
    f3 = TN.flat_params(m.parameters(), owner=m)
    assert f3 is not f[0]                                   # the backward pass consumed the shared node
    copy.deepcopy(m)                                        # nothing graph-attached lives in the module
    with torch.no_grad():
        assert not TN.flat_params(m.parameters(), owner=m).requires_grad
        m[0].weight.add_(1.0)                               # what an optimizer step does
    f4 = TN.flat_params(m.parameters(), owner=m)
    assert f4 is not f3 and torch.equal(f4[:12], m[0].weight.reshape(-1))
