"""GPU checks at BASELINE.json's full sizes (configs[1]: YOLOv8n@640, batch 64; one 160x160 map of
configs[2]) through size-independent properties -- the oracle only finishes small cases in seconds:

* K1: sums against fp64 torch sums (tolerance), per-channel min / max exact;
* K3: integer codes inside the assigned width's range, y == (code - zp) * scale * m recomputed
  element-wise with torch (bit exact in fp32, bf16 after rounding), idempotence Q(Q(x)) == Q(x),
  monotonicity in x within (channel, bit width);
* whole hook: bit maps integral in [2, 8], mask in (0, 1], determinism, batch-composition
  independence of everything per-image (bit map of image i does not depend on the other images),
  sharded (two virtual ranks) == unsharded.
"""
import pytest
import torch

pytestmark = pytest.mark.gpu

SHAPES = [(64, 80, 80), (128, 40, 40), (256, 20, 20)]


@pytest.fixture(scope="module")
def env():
    from golden_util import weights
    from mcaq_yolo_b200 import fused, modules as M, ops
    W = weights()
    a, m, _ = M.build_fixture_modules(W, "cuda")
    qs = [M.build_fixture_modules(W, "cuda")[2] for _ in SHAPES]
    return dict(M=M, ops=ops, fused=fused, a=a, m=m, qs=qs, W=W)


def synth(B, C, H, Wd, dtype, seed):
    g = torch.Generator(device="cuda")
    g.manual_seed(seed)
    coarse = torch.randn(B, C, H // 8 + 2, Wd // 8 + 2, device="cuda", generator=g)
    up = torch.nn.functional.interpolate(coarse, size=(H, Wd), mode="bilinear", align_corners=False)
    bias = torch.randn(1, C, 1, 1, device="cuda", generator=g) * 0.5
    return (up * 1.6 + 0.1 * torch.randn(B, C, H, Wd, device="cuda", generator=g) + bias + 0.3).to(dtype).contiguous()


@pytest.mark.parametrize("shape", SHAPES + [(128, 160, 160)])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_k1_full_size(shape, dtype, env):
    ops = env["ops"]
    C, H, Wd = shape
    B = 64 if H < 160 else 8
    x = synth(B, C, H, Wd, dtype, 3)
    s, a, keys = ops.reduce_planes(x)
    packed = ops.ranges_decode(keys)
    xf = x.double()
    torch.testing.assert_close(s.double(), xf.sum(1), rtol=0, atol=2e-5 * C)
    torch.testing.assert_close(a.double(), xf.abs().sum(1), rtol=1e-6, atol=2e-5 * C)
    assert torch.equal(packed[:C], x.float().amin(dim=(0, 2, 3)))
    assert torch.equal(-packed[C:], x.float().amax(dim=(0, 2, 3)))


@pytest.mark.parametrize("shape", SHAPES + [(128, 160, 160)])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_k3_full_size_properties(shape, dtype, env):
    ops = env["ops"]
    C, H, Wd = shape
    B = 64 if H < 160 else 8
    x = synth(B, C, H, Wd, dtype, 5)
    tile = ops.tile_size(H, 8)
    Ht, Wt = H // tile, Wd // tile
    g = torch.Generator(device="cuda")
    g.manual_seed(11)
    bits = torch.randint(2, 9, (B, Ht, Wt), device="cuda", generator=g).float()
    m = torch.rand(B, H, Wd, device="cuda", generator=g) * 0.3 + 0.7
    _, _, keys = ops.reduce_planes(x)
    packed = ops.ranges_decode(keys)
    qt = ops.build_qtable(packed)                                    # (7, C, 2) scale / zero point
    y, codes = ops.tile_quantize(x, bits, qt, m, want_codes=True)
    y2 = ops.tile_quantize_ranges(x, bits, packed, None, None, m)
    assert torch.equal(y, y2), "table kernel and ranges kernel disagree"
    # per-element bit width, scale, zero point
    bpix = bits.repeat_interleave(tile, 1).repeat_interleave(tile, 2)          # (B,H,W)
    bidx = (bpix.long() - 2)[:, None].expand(B, C, H, Wd)
    cidx = torch.arange(C, device="cuda")[None, :, None, None].expand(B, C, H, Wd)
    scale = qt[bidx, cidx, 0]
    zp = qt[bidx, cidx, 1]
    qmax = (2.0 ** (bpix - 1) - 1)[:, None]
    qmin = (-(2.0 ** (bpix - 1)))[:, None]
    cf = codes.float()
    assert bool(((cf >= qmin) & (cf <= qmax)).all()), "code outside the assigned width"
    want = ((cf - zp) * scale) * m[:, None]
    assert torch.equal(want.to(dtype), y), "y != (code - zp) * scale * m"
    # the codes are the rounded affine image of x (independent torch restatement)
    code_t = torch.clamp(torch.round(x.float() / scale + zp), qmin, qmax)
    assert torch.equal(code_t, cf)
    if dtype == torch.float32:
        # idempotence without the mask: quantization levels are fixed points
        y0 = ops.tile_quantize_ranges(x, bits, packed, None, None, None)
        y1 = ops.tile_quantize_ranges(y0, bits, packed, None, None, None)
        assert torch.equal(y0, y1)
    # monotone in x for a fixed (image, channel, tile): sort one tile's pixels
    xs = x[0, 0, :tile, :tile].float().reshape(-1)
    ys = ops.tile_quantize_ranges(x, bits, packed, None, None, None)[0, 0, :tile, :tile].float().reshape(-1)
    order = torch.argsort(xs)
    assert bool((ys[order][1:] >= ys[order][:-1]).all())


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_hook_full_size_properties(dtype, env):
    fused, a, m, qs = env["fused"], env["a"], env["m"], env["qs"]
    feats = [synth(64, C, H, Wd, dtype, 20 + i) for i, (C, H, Wd) in enumerate(SHAPES)]
    hot = fused.FusedHotPath(a, m, qs, streams=True)
    r1 = hot.run(feats)
    r2 = hot.run(feats)
    torch.cuda.synchronize()
    for u, v, x in zip(r1, r2, feats):
        assert torch.equal(u["features_q"], v["features_q"]) and torch.equal(u["bit_map"], v["bit_map"]), "not deterministic"
        b = u["bit_map"]
        assert bool(((b >= 2) & (b <= 8) & (b == b.round())).all())
        assert bool(((u["complexity"] >= 0) & (u["complexity"] <= 1)).all())
        assert u["features_q"].shape == x.shape and u["features_q"].dtype == x.dtype
        assert bool(torch.isfinite(u["features_q"].float()).all())
    # per-image quantities do not depend on the rest of the batch: reverse the batch order
    rev = hot.run([f.flip(0).contiguous() for f in feats])
    torch.cuda.synchronize()
    for u, v in zip(r1, rev):
        assert torch.equal(u["bit_map"], v["bit_map"].flip(0))
        assert torch.equal(u["complexity"], v["complexity"].flip(0))
        assert torch.equal(u["features_q"], v["features_q"].flip(0)), "ranges are batch-wide: order must not matter"


def test_sharded_full_size(env):
    """Two virtual ranks of 32 images each == one batch of 64 (range merge over the exchange)."""
    fused, a, m, qs = env["fused"], env["a"], env["m"], env["qs"]
    from mcaq_yolo_b200.peer import RangeExchange
    C, H, Wd = SHAPES[0]
    x = synth(64, C, H, Wd, torch.bfloat16, 31)
    x[:32] *= 1.5                                    # the two shards see different ranges
    with torch.no_grad():
        ref, _ = fused.fused_scale_forward(x, a, m, qs[0], 1.0, None)
    torch.cuda.synchronize()
    ex = RangeExchange.virtual(C, 2)
    streams = [torch.cuda.Stream(), torch.cuda.Stream()]
    shards = [x[:32].contiguous(), x[32:].contiguous()]
    recs = [None, None]
    torch.cuda.synchronize()
    for r in range(2):
        with torch.cuda.stream(streams[r]), torch.no_grad():
            recs[r], _ = fused.fused_scale_forward(shards[r], a, m, qs[0], 1.0, None, xchg=ex[r])
    torch.cuda.synchronize()
    assert torch.equal(torch.cat([r_["features_q"] for r_ in recs], 0), ref["features_q"])
    assert torch.equal(torch.cat([r_["bit_map"] for r_ in recs], 0), ref["bit_map"])
