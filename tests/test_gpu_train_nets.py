"""Native training kernels of the three tile-level networks (csrc/train_nets.cu) against torch autograd
restatements with the reference's semantics (tests/torch_nets_ref.py): outputs, input gradients and every
parameter gradient; BatchNorm running statistics; the in-kernel SyncBN merge with two virtual ranks."""
import copy

import numpy as np
import pytest
import torch

import torch_nets_ref as R
from golden_util import Case, weights

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def M():
    from mcaq_yolo_b200 import modules
    return modules


@pytest.fixture(scope="module")
def W():
    return weights()


def close(a, b, rtol=2e-3, atol_rel=2e-4, what="", scale=None):
    a, b = a.detach().float().cpu().numpy(), b.detach().float().cpu().numpy()
    scale = max(1e-12, np.abs(b).max()) if scale is None else scale
    np.testing.assert_allclose(a, b, rtol=rtol, atol=atol_rel * scale, err_msg=what)


def grads_close(mod, ref, rtol=2e-3, atol_rel=2e-4, combine=None):
    """Every parameter gradient of `mod` against `ref`; the absolute tolerance is relative to the LARGEST gradient
    of the network: a Linear bias feeding a BatchNorm has a mathematically zero gradient (both sides hold rounding
    noise of ~1e-10 there)."""
    scale = max(float(p.grad.abs().max()) for p in ref.parameters())
    for (n, p), (_, pr) in zip(mod.named_parameters(), ref.named_parameters()):
        assert p.grad is not None, n
        close(p.grad if combine is None else combine(n), pr.grad, rtol=rtol, atol_rel=atol_rel, what=f"grad {n}", scale=scale)


@pytest.mark.parametrize("name", ["small_smooth", "c3_v8n_smooth", "c5_v8n_smooth"])
def test_complexity_training_backward(name, M, W):
    torch.backends.cuda.matmul.allow_tf32 = False
    c = Case(name)
    x = torch.from_numpy(c.x()).cuda()
    a, _, _ = M.build_fixture_modules(W, "cuda", grid_size=c.grid)
    a.train()
    ref = copy.deepcopy(a)
    phi, _ = a.compute_phi_tiles(x)
    torch.manual_seed(0)
    g = torch.randn(phi.shape[:3], device="cuda")
    out = a(x)
    out.backward(g)
    out_r = R.complexity(ref, phi)
    out_r.backward(g)
    close(out, out_r, rtol=1e-4, atol_rel=1e-5, what="complexity forward")
    grads_close(a.complexity_mlp, ref.complexity_mlp)


@pytest.fixture
def mapper_cluster():
    """Force the cluster size of the mapper kernels for one test (0 = the probed default), restored afterwards."""
    from mcaq_yolo_b200 import _lib
    lib = _lib.load()
    yield lib.mcaq_debug_mapper_cluster
    lib.mcaq_debug_mapper_cluster(0)


# rows per CTA <= 218 keep the per-row records in shared memory, more fall back to the global scratch:
# (16,10,10): 100 / 200 rows per CTA (cached either way); (40,10,10): 250 (16 CTAs) / 500 (8 CTAs): global path
@pytest.mark.parametrize("shape,temp,cluster", [((2, 5, 5), 1.0, 0), ((4, 10, 10), 1.3, 0), ((16, 10, 10), 0.7, 0),
                                                ((3, 7, 9), None, 0), ((16, 10, 10), 1.0, 8), ((16, 10, 10), 1.0, 16),
                                                ((40, 10, 10), 1.0, 16), ((40, 10, 10), 0.9, 8), ((25, 10, 10), 1.0, 8)])
def test_mapper_training_forward_backward_and_running_stats(shape, temp, cluster, M, W, mapper_cluster):
    torch.backends.cuda.matmul.allow_tf32 = False
    mapper_cluster(cluster)
    _, m, _ = M.build_fixture_modules(W, "cuda")
    m.train()
    ref = copy.deepcopy(m)
    torch.manual_seed(1)
    c = (torch.rand(shape, device="cuda") * 1.2 - 0.1).requires_grad_(True)     # some values outside [0, 1]
    cr = c.detach().clone().requires_grad_(True)
    g = torch.randn(shape, device="cuda")
    bits = m(c, temp, return_continuous=True)
    bits.backward(g)
    bits_r = R.mapper(ref, cr, temp)
    bits_r.backward(g)
    close(bits, bits_r, rtol=1e-4, atol_rel=1e-5, what="bits")
    close(c.grad, cr.grad, what="grad complexity")
    grads_close(m.mapping_network, ref.mapping_network)
    for i in (1, 4, 7):
        bn, bnr = m.mapping_network[i], ref.mapping_network[i]
        close(bn.running_mean, bnr.running_mean, rtol=1e-4, atol_rel=1e-5, what=f"running_mean {i}")
        close(bn.running_var, bnr.running_var, rtol=1e-4, atol_rel=1e-5, what=f"running_var {i}")
        assert int(bn.num_batches_tracked) == int(bnr.num_batches_tracked)
    # rounded output keeps the straight-through gradient
    b2 = m(c.detach().requires_grad_(True), temp, return_continuous=False)
    assert torch.equal(b2, torch.round(b2))


@pytest.mark.parametrize("name", ["small_smooth", "c3_v8n_smooth", "crop_50"])
def test_soft_mask_training_backward(name, M, W):
    c = Case(name)
    x = torch.from_numpy(c.x()).cuda()
    _, _, q = M.build_fixture_modules(W, "cuda")
    sm = q.soft_mask.train()
    ref = copy.deepcopy(sm)
    bf = torch.from_numpy(c["bit_map_frac"]).cuda()
    b1 = bf.clone().requires_grad_(True)
    b2 = bf.clone().requires_grad_(True)
    torch.manual_seed(2)
    g = torch.randn(c.B, 1, c.H, c.W, device="cuda")
    m = sm(b1, x)
    m.backward(g)
    mr = R.soft_mask(ref, b2, x)
    mr.backward(g)
    close(m, mr, rtol=1e-4, atol_rel=1e-5, what="mask")
    close(b1.grad, b2.grad, rtol=5e-3, atol_rel=2e-3, what="grad bit_map")       # fp32 atomics vs cuDNN's order
    grads_close(sm.net, ref.net, rtol=5e-3, atol_rel=2e-3)


def test_bit_map_losses(M):
    from mcaq_yolo_b200 import train_nets as TN
    torch.manual_seed(3)
    maps = [(torch.rand(4, 10, 10, device="cuda") * 6 + 2).requires_grad_(True),
            (torch.rand(4, 5, 5, device="cuda") * 6 + 2).requires_grad_(True)]
    refs = [m.detach().clone().requires_grad_(True) for m in maps]
    avg, lbit, lsm = TN.bit_map_losses(maps, 4.0)
    (0.3 * lbit + 0.7 * lsm + 0.1 * avg).backward()
    avg_r = torch.stack([m.float().mean() for m in refs]).mean()          # models/mcaq_yolo.py:575
    tv = []
    for m in refs:                                                         # models/mcaq_yolo.py:86-108
        dx = (m[:, 1:, :] - m[:, :-1, :]).abs()
        dy = (m[:, :, 1:] - m[:, :, :-1]).abs()
        tv.append((dx.sum() + dy.sum()) / (dx.numel() + dy.numel()))
    lsm_r = sum(tv) / len(tv)
    (0.3 * (avg_r - 4.0) ** 2 + 0.7 * lsm_r + 0.1 * avg_r).backward()
    close(avg, avg_r, rtol=1e-5, atol_rel=1e-6)
    close(lsm, lsm_r, rtol=1e-5, atol_rel=1e-6)
    for a, b in zip(maps, refs):
        close(a.grad, b.grad, rtol=1e-4, atol_rel=1e-5)


def test_mapper_syncbn_two_virtual_ranks(M, W):
    """Two ranks' mapper kernels on two streams of ONE GPU, statistics merged through the peer exchange inside the
    kernels: every rank's bits / gradients equal those of the unsharded batch, running statistics too."""
    from mcaq_yolo_b200.peer import RangeExchange
    torch.backends.cuda.matmul.allow_tf32 = False
    _, m0, _ = M.build_fixture_modules(W, "cuda")
    m0.train()
    ranks = [copy.deepcopy(m0) for _ in range(2)]
    full = copy.deepcopy(m0)
    ex = RangeExchange.virtual(128, 2)
    for r, mm in enumerate(ranks):
        mm.stat_exchange = ex[r]
    torch.manual_seed(4)
    c = torch.rand(6, 10, 10, device="cuda")
    g = torch.randn(6, 10, 10, device="cuda")
    parts = [(c[:4].clone().requires_grad_(True), g[:4]), (c[4:].clone().requires_grad_(True), g[4:])]   # uneven shards
    streams = [torch.cuda.Stream(), torch.cuda.Stream()]
    outs = [None, None]
    torch.cuda.synchronize()
    for r in range(2):
        with torch.cuda.stream(streams[r]):
            outs[r] = ranks[r](parts[r][0], 1.0, return_continuous=True)
    torch.cuda.synchronize()
    for r in range(2):
        ex[r].check()
    for r in range(2):
        with torch.cuda.stream(streams[r]):
            outs[r].backward(parts[r][1])
    torch.cuda.synchronize()
    for r in range(2):
        ex[r].check()
    cf = c.clone().requires_grad_(True)
    bf = R.mapper(full, cf, 1.0)
    bf.backward(g)
    close(torch.cat(outs), bf, rtol=1e-4, atol_rel=1e-5, what="bits of the sharded batch")
    close(torch.cat([parts[0][0].grad, parts[1][0].grad]), cf.grad, what="grad complexity")
    g1 = dict(ranks[1].mapping_network.named_parameters())
    # the flat gradient all-reduce adds the shards' parameter gradients
    grads_close(ranks[0].mapping_network, full.mapping_network,
                combine=lambda n: dict(ranks[0].mapping_network.named_parameters())[n].grad + g1[n].grad)
    for i in (1, 4, 7):
        for mm in ranks:
            close(mm.mapping_network[i].running_var, full.mapping_network[i].running_var, rtol=1e-4, atol_rel=1e-5)
            close(mm.mapping_network[i].running_mean, full.mapping_network[i].running_mean, rtol=1e-4, atol_rel=1e-5)


def test_whole_training_hook_matches_torch_nets(M, W):
    """One train-mode hook (analyzer -> mapper -> quantiser with soft mask) end to end: gradients of every small
    network against the same hook with the torch restatements of the three networks."""
    from mcaq_yolo_b200 import ops
    torch.backends.cuda.matmul.allow_tf32 = False
    c = Case("small_smooth")
    x = torch.from_numpy(c.x()).cuda()
    g = torch.from_numpy(c.grad()).cuda()
    a, m, q = M.build_fixture_modules(W, "cuda")
    for mod in (a, m, q):
        mod.train()
    ar, mr, qr = copy.deepcopy(a), copy.deepcopy(m), copy.deepcopy(q)
    rec = M.mcaq_hook_forward(x, a, m, q, temperature=1.0, training=True)
    (rec["features_q"] * g).sum().backward()
    # comparator: same kernels for the HBM sweeps and the fractional quantiser, torch for the three networks
    phi, _ = ar.compute_phi_tiles(x)
    cpx = R.complexity(ar, phi)
    bits = R.mapper(mr, cpx, 1.0)
    qr.update_running_stats(x)
    qt = ops.build_qtable(None, qr.running_min, qr.running_max)
    mask = R.soft_mask(qr.soft_mask, bits, x)
    y = M._FractionalQuant.apply(x, bits.float(), mask, qt)
    (y * g).sum().backward()
    close(rec["bit_map"], bits, rtol=1e-4, atol_rel=1e-5, what="continuous bit map")
    close(rec["features_q"], y, rtol=1e-3, atol_rel=1e-4, what="quantised features")
    for mod, ref in ((a.complexity_mlp, ar.complexity_mlp), (m.mapping_network, mr.mapping_network), (q.soft_mask.net, qr.soft_mask.net)):
        grads_close(mod, ref, rtol=2e-2, atol_rel=5e-3)
