"""The plane pipeline (csrc/morph_planes.cu): phi tiles of image-sized inputs -- the reference's curriculum
scoring path (utils/dataset.py:345-353 -> core/morphology.py:923-937) -- against tests/golden/r2/image_*.npz
(written by tools/make_golden_r2.py from the RUNNING reference) and the oracle, and cross-checked against the
fused per-image kernel on every feature-map sized golden case (the two implementations must agree bit for bit
on planes and integer counts)."""
import hashlib
import os

import numpy as np
import pytest
import torch

import mcaq_oracle as o
from golden_util import CASE_NAMES, GOLDEN_DIR, Case, weights
from inputs import feature_map

pytestmark = pytest.mark.gpu
RTOL, ATOL = 1e-4, 2e-6


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def unpack_bits(words, Wc):
    w = words.cpu().numpy().astype(np.uint32)
    bits = ((w[..., None] >> np.arange(32, dtype=np.uint32)) & 1).astype(bool)
    return bits.reshape(w.shape[0], w.shape[1], -1)[:, :, :Wc]


@pytest.fixture(scope="module")
def env():
    from mcaq_yolo_b200 import constants as K, modules as M, ops
    return dict(K=K, M=M, ops=ops, W=weights())


IMAGES = [("image_320", 320, 51, "smooth"), ("image_640", 640, 52, "smooth"), ("image_640_noise", 640, 53, "noise"),
          ("image_1280", 1280, 54, "smooth")]


@pytest.mark.parametrize("name,S,seed,kind", IMAGES)
def test_image_sized_phi_matches_reference(name, S, seed, kind, env):
    ops, K = env["ops"], env["K"]
    g = np.load(os.path.join(GOLDEN_DIR, "r2", name + ".npz"))
    tile, ht = int(g["cfg"][6]), int(g["cfg"][7])
    x = torch.from_numpy(feature_map(kind, 1, 3, S, S, seed)).cuda()
    s, _, _ = ops.reduce_planes(x)
    assert not ops.morph_fits(1, 3, S, S, 8)
    phi, dbg = ops.morph_phi(s, 3, 8, K.device_constants("cuda"), debug=True)
    Hc = ht * tile
    edge = unpack_bits(dbg["edge_bits"], Hc)
    binm = unpack_bits(dbg["bin_bits"], Hc)
    assert int(edge.sum()) == int(g["edge_count"]) and int(binm.sum()) == int(g["bin_count"])
    assert sha(np.packbits(edge[0])) == str(g["edge_sha"]), "Canny edge plane differs from the reference"
    assert sha(np.packbits(binm[0])) == str(g["bin_sha"]), "adaptive-threshold plane differs from the reference"
    np.testing.assert_allclose(dbg["gray"].cpu().numpy()[0, ::7, ::5], g["gray_sub"], rtol=0, atol=0)
    p = phi.cpu().numpy()
    np.testing.assert_allclose(p, g["phi"], rtol=RTOL, atol=ATOL)
    for k in (3, 4):                                    # integer-count metrics: exact
        assert np.array_equal(p[..., k], g["phi"][..., k])
    # the analyzer API on the raw image (what compute_dataset_complexity calls per image)
    M = env["M"]
    a, _, _ = M.build_fixture_modules(env["W"], "cuda")
    with torch.no_grad():
        sc = a.score_image(x)
        cpx = a(x)
    np.testing.assert_allclose(sc.cpu().numpy(), g["score"], rtol=RTOL, atol=ATOL)
    np.testing.assert_allclose(cpx.cpu().numpy(), g["complexity"], rtol=RTOL, atol=5e-6)


def test_image_640_integer_counts_against_oracle(env):
    ops, K = env["ops"], env["K"]
    x = feature_map("smooth", 1, 3, 640, 640, 52)
    d = {}
    phi_o = o.phi_tiles(x, 8, detail=d)
    s, _, _ = ops.reduce_planes(torch.from_numpy(x).cuda())
    phi, dbg = ops.morph_phi(s, 3, 8, K.device_constants("cuda"), debug=True)
    cnt = dbg["counts"].cpu().numpy()
    assert np.array_equal(dbg["gray"].cpu().numpy(), d["gray"])
    assert np.array_equal(cnt[..., 11][:, 0, 0], d["otsu_bin"])
    assert np.array_equal(cnt[..., 0], d["edge_count"])
    assert np.array_equal(cnt[..., 1], d["area"]) and np.array_equal(cnt[..., 2], d["perim"])
    assert np.array_equal(cnt[..., 3], d["euler_x4"])
    S = d["box_counts"].shape[0]
    assert S == 6 and np.array_equal(cnt[..., 4:4 + S], d["box_counts"].transpose(1, 2, 3, 0))
    assert np.array_equal(dbg["lbp_hist"].cpu().numpy(), d["lbp_hist"].transpose(0, 2, 3, 1))
    np.testing.assert_allclose(phi.cpu().numpy(), phi_o, rtol=RTOL, atol=ATOL)


@pytest.mark.parametrize("name", CASE_NAMES)
def test_plane_pipeline_equals_fused_kernel(name, env):
    """Feature-map sized golden cases through BOTH implementations: tiles of 4..32 pixels inside 32x32 regions,
    cropped and rectangular planes, batch > 1."""
    ops, K = env["ops"], env["K"]
    c = Case(name)
    s, _, _ = ops.reduce_planes(torch.from_numpy(c.x()).cuda())
    consts = K.device_constants("cuda")
    phi_f, df = ops.morph_phi(s, c.C, c.grid, consts, debug=True)
    phi_p, dp = ops.morph_phi(s, c.C, c.grid, consts, debug=True, force_planes=True)
    assert torch.equal(df["gray"], dp["gray"])
    assert torch.equal(df["edge_bits"], dp["edge_bits"]) and torch.equal(df["bin_bits"], dp["bin_bits"])
    assert torch.equal(df["lbp_hist"], dp["lbp_hist"])
    cf, cp = df["counts"].cpu().numpy(), dp["counts"].cpu().numpy()
    assert np.array_equal(cf[..., :4], cp[..., :4])
    S = int(np.log2(c.tile))
    assert np.array_equal(cf[..., 4:4 + S], cp[..., 4:4 + S]) and np.array_equal(cf[..., 9], cp[..., 11])
    # phi: integer-count metrics identical, the rest within the contract (log tables vs fp64 log at run time)
    pf, pp = phi_f.cpu().numpy(), phi_p.cpu().numpy()
    for k in (3, 4):
        assert np.array_equal(pf[..., k], pp[..., k])
    np.testing.assert_allclose(pp, pf, rtol=1e-6, atol=1e-7)
    np.testing.assert_allclose(pp, c["phi"], rtol=RTOL, atol=ATOL)


def test_batched_scoring_is_per_image(env):
    """score_image over a batch equals scoring each image alone (the reference loops at batch 1)."""
    M = env["M"]
    a, _, _ = M.build_fixture_modules(env["W"], "cuda")
    x = torch.from_numpy(np.concatenate([feature_map("smooth", 1, 3, 640, 640, 60 + i) for i in range(3)])).cuda()
    with torch.no_grad():
        sb = a.score_image(x)
        ss = torch.cat([a.score_image(x[i:i + 1]) for i in range(3)])
    assert torch.equal(sb, ss)
