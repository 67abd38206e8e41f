#!/usr/bin/env python
"""Generate tests/golden/*.npz by RUNNING THE REAL REFERENCE (imported read-only from
/root/reference) on the deterministic inputs of tests/inputs.py.

Run in the build container only:   python tools/make_golden.py
The GPU box has no /root/reference; tests there read the committed .npz files.

Stored per case: the reference's outputs for every stage of the hot path
(SURVEY.md 8c): normalised gray plane, Canny edge / strong-free binary planes, adaptive
binary mask, phi(8), complexity, bit maps (MLP + linear mapper), soft mask, per-channel
ranges, a strided subsample + sha256 of the de-quantised map and of the integer codes,
and the training-path forward/backward.  Weights of the fixtures are in weights.npz.
"""
import hashlib
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from ref_loader import load_reference  # noqa: E402
from inputs import feature_map, fractional_bit_map, integer_bit_map  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")

# (name, kind, B, C, H, W, grid, seed)
CASES = [
    ("parity_2x8x32", "noise", 2, 8, 32, 32, 8, 11),          # tests/test_smoke.py:239 shape
    ("small_noise", "noise", 2, 16, 40, 40, 8, 12),
    ("small_smooth", "smooth", 2, 16, 40, 40, 8, 13),
    ("c3_v8n_smooth", "smooth", 2, 64, 80, 80, 8, 14),
    ("c3_v8n_noise", "noise", 1, 64, 80, 80, 8, 15),
    ("c4_v8n_smooth", "smooth", 2, 128, 40, 40, 8, 16),
    ("c5_v8n_smooth", "smooth", 2, 256, 20, 20, 8, 17),
    ("c3_v8s1280_smooth", "smooth", 1, 128, 160, 160, 8, 18),
    ("c3_grid4", "smooth", 1, 64, 80, 80, 4, 19),
    ("c3_grid16", "smooth", 1, 64, 80, 80, 16, 20),
    ("crop_50", "smooth", 1, 8, 50, 50, 8, 21),               # H % tile != 0 -> crop + nearest
    ("rect_24x40", "smooth", 2, 12, 24, 40, 8, 22),           # non-square
]


def sha(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def npw(module):
    return {k: v.detach().cpu().numpy().copy() for k, v in module.state_dict().items()}


def build_fixtures(morph, ba, qz):
    """SURVEY 8c fixtures: perturbed analyzer, spread mapper with non-trivial BN
    running stats, soft-mask net with visible spatial variation."""
    torch.manual_seed(1234)
    A = morph.MorphologicalComplexityAnalyzer(grid_size=8, device="cpu")
    with torch.no_grad():
        for i in (1, 4):
            A.complexity_mlp[i].weight.add_(0.1 * torch.randn_like(A.complexity_mlp[i].weight))
            A.complexity_mlp[i].bias.add_(0.1 * torch.randn_like(A.complexity_mlp[i].bias))
        A.feature_weights.copy_(torch.tensor([0.3, 0.1, 0.25, 0.15, 0.2]))
    A.eval()
    # centre/spread the analyzer's output on real phi so C covers ~[0.1, 0.9]
    with torch.no_grad():
        phis = []
        for kind, C, H, seed in (("smooth", 64, 80, 1), ("noise", 64, 80, 2), ("smooth", 128, 40, 3),
                                 ("smooth", 256, 20, 4), ("noise", 16, 40, 5)):
            p, _ = A.compute_phi_tiles(torch.from_numpy(feature_map(kind, 2, C, H, H, seed)))
            phis.append(p.reshape(-1, 8))
        phis = torch.cat(phis)
        z = A.complexity_mlp[:-1](phis)
        s = 1.3 / z.std()
        A.complexity_mlp[6].weight.mul_(s)
        A.complexity_mlp[6].bias.mul_(s).sub_(s * z.mean())
        cvals = A.complexity_mlp(phis).reshape(-1)

    M = ba.ComplexityToBitMappingNetwork()
    with torch.no_grad():
        # a fresh mapper is dead below the batch-mean complexity (all-positive weights,
        # BN centring, ReLU): give the BN affines trained-looking positive values
        for i in (1, 4, 7):
            M.mapping_network[i].weight.uniform_(0.5, 1.5)
            M.mapping_network[i].bias.uniform_(0.5, 2.0)
    M.train()
    for _ in range(6):
        M(cvals[torch.randperm(cvals.numel())[:400]].reshape(4, 10, 10))
    M.eval()
    with torch.no_grad():
        z = M.mapping_network[:-1](M.create_augmented_features(cvals.reshape(-1, 1)))
        s = 2.5 / z.std()
        M.mapping_network[9].weight.mul_(s)
        M.mapping_network[9].bias.mul_(s).sub_(s * z.mean())

    Q = qz.SpatialAdaptiveQuantization(calibration_mode="minmax", smooth_transitions=True,
                                       per_channel=True)
    with torch.no_grad():
        Q.soft_mask.net[0].weight.normal_(0, 0.6)
        Q.soft_mask.net[0].bias.normal_(0, 0.2)
        Q.soft_mask.net[2].weight.normal_(0, 0.8)
        Q.soft_mask.net[2].bias.copy_(torch.tensor([1.0, 0.0]))
    Q.eval()
    return A, M, Q


def ref_codes(Q, x, bit_map):
    """Integer codes the reference computes internally (quantization.py:597-600),
    recomputed with its own get_calibration_params, per SURVEY 8c."""
    B, C, H, W = x.shape
    codes = torch.zeros_like(x)
    import torch.nn.functional as F
    for bv in torch.unique(bit_map):
        bits = int(round(float(bv)))
        scale, zp = Q.get_calibration_params(x, bits)
        qp = sys.modules[Q.__module__].QuantizationParameters(bits)
        q = torch.clamp(torch.round(x / scale + zp), qp.qmin, qp.qmax)
        sel = F.interpolate((bit_map == bv).float().unsqueeze(1), size=(H, W), mode="nearest")
        codes = codes + q * sel
    return codes.to(torch.int16).numpy()


def run_case(name, kind, B, C, H, W, grid, seed, morph, ba, qz, A, M, Q):
    A.grid_size = grid
    x_np = feature_map(kind, B, C, H, W, seed)
    x = torch.from_numpy(x_np)
    out = {}
    tile = A._tile_size(H)
    ht, wt = H // tile, W // tile
    Hc, Wc = ht * tile, wt * tile
    with torch.no_grad():
        gray = A._normalize01(x[:, :, :Hc, :Wc].mean(dim=1, keepdim=True).float())
        edge = A._gpu_canny(gray)
        binm = A._binarize(gray)
        phi, _ = A.compute_phi_tiles(x)
        raw = A.complexity_mlp(phi.reshape(-1, 8)).reshape(B, ht, wt)
        cpx = A(x)
        bm_mlp = M(cpx, 1.0)
        bm_mlp_c = M(cpx, 1.3, return_continuous=True)
        L = ba.LinearBitMapper()
        bm_lin = L(cpx, 1.0)
        score = A.score_image(x)
        # quantizer (eval, dynamic ranges), MLP bit map
        Q.eval()
        Q.stats_frozen = torch.tensor(False)
        Q.running_min = None
        Q.running_max = None
        m = Q.soft_mask(bm_mlp, x)
        y = Q(x, bm_mlp, training=False)
        codes = ref_codes(Q, x, bm_mlp)
        # same with a uniformly random integer map (all widths 2..8 present)
        bm_rand = torch.from_numpy(integer_bit_map(B, ht, wt, seed))
        y_rand = Q(x, bm_rand, training=False)
        codes_rand = ref_codes(Q, x, bm_rand)
        mn = x.amin(dim=(0, 2, 3)).numpy()
        mx = x.amax(dim=(0, 2, 3)).numpy()

    out.update(
        cfg=np.array([B, C, H, W, grid, seed, tile, ht, wt], dtype=np.int64),
        kind=np.array(kind),
        gray=gray[:, 0].numpy(),
        edge=np.packbits(edge[:, 0].numpy() > 0),
        binmask=np.packbits(binm[:, 0].numpy() > 0),
        phi=phi.numpy(), complexity_raw=raw.numpy(), complexity=cpx.numpy(),
        bit_map_mlp=bm_mlp.numpy(), bit_map_mlp_cont_T13=bm_mlp_c.numpy(),
        bit_map_linear=bm_lin.numpy(), score=score.numpy(),
        soft_mask=m[:, 0].numpy(), ch_min=mn, ch_max=mx,
        y_sub=y.numpy()[:, ::3, ::5, ::7], y_sha=np.array(sha(y.numpy())),
        codes_sha=np.array(sha(codes)), codes_sub=codes[:, ::3, ::5, ::7],
        bit_map_rand=bm_rand.numpy(),
        y_rand_sub=y_rand.numpy()[:, ::3, ::5, ::7], y_rand_sha=np.array(sha(y_rand.numpy())),
        codes_rand_sha=np.array(sha(codes_rand)),
    )

    # training path: EMA stats adopt the batch statistics on the first call
    # (quantization.py:342-344), continuous bit map, soft mask on.
    Q.train()
    Q.running_min = None
    Q.running_max = None
    bf = torch.from_numpy(fractional_bit_map(B, ht, wt, seed)).requires_grad_(True)
    xg = x.clone().requires_grad_(True)
    g = torch.from_numpy(feature_map("noise", B, C, H, W, seed + 500))
    yt = Q(xg, bf, training=True)
    yt.backward(g)
    with torch.no_grad():
        m_t = Q.soft_mask(bf.detach(), x)
    out.update(
        bit_map_frac=bf.detach().numpy(),
        train_y_sha=np.array(sha(yt.detach().numpy())),
        train_y_sub=yt.detach().numpy()[:, ::3, ::5, ::7],
        train_dx_sha=np.array(sha(xg.grad.numpy())),
        train_dx_sub=xg.grad.numpy()[:, ::3, ::5, ::7],
        train_dbit_total=bf.grad.numpy(),           # fractional path + soft-mask path
        train_soft_mask=m_t[:, 0].numpy(),
        train_run_min=Q.running_min.numpy().ravel(), train_run_max=Q.running_max.numpy().ravel(),
    )
    # mask-off training backward isolates the fractional-bit term (test_smoke.py:103-112)
    Q2 = qz.SpatialAdaptiveQuantization(smooth_transitions=False)
    Q2.train()
    bf2 = torch.from_numpy(fractional_bit_map(B, ht, wt, seed)).requires_grad_(True)
    yt2 = Q2(x, bf2, training=True)
    yt2.backward(g)
    out.update(train_nomask_y_sha=np.array(sha(yt2.detach().numpy())),
               train_nomask_dbit=bf2.grad.numpy())
    # second EMA step on a shifted batch (quantization.py:346-347)
    with torch.no_grad():
        Q.train()
        Q(x * 1.5 + 0.25, bf.detach(), training=True)
        out.update(ema2_min=Q.running_min.numpy().ravel(), ema2_max=Q.running_max.numpy().ravel())
    Q.eval()
    Q.running_min = None
    Q.running_max = None
    np.savez_compressed(os.path.join(OUT, f"{name}.npz"), **out)
    return out


def main():
    morph, ba, qz = load_reference()
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(8)      # H*W/8 threads keeps torch's vectorised sum path on all pixels
    A, M, Q = build_fixtures(morph, ba, qz)
    w = {}
    for prefix, mod in (("analyzer.", A), ("mapper.", M), ("quantizer.", Q)):
        for k, v in npw(mod).items():
            w[prefix + k] = v
    # fixed stencils as the reference computes them (torch), to pin the oracle's constants
    x1 = torch.arange(5, dtype=torch.float32) - 2
    g1 = torch.exp(-(x1 ** 2) / 2.0)
    g1 = g1 / g1.sum()
    w["const.canny_blur"] = (g1.unsqueeze(0) * g1.unsqueeze(1)).numpy()
    k = 11
    sigma = 0.3 * ((k - 1) * 0.5 - 1) + 0.8
    xa = torch.arange(k, dtype=torch.float32) - k // 2
    ga = torch.exp(-(xa ** 2) / (2 * sigma ** 2))
    ga = ga / ga.sum()
    w["const.adapt_blur"] = (ga.unsqueeze(0) * ga.unsqueeze(1)).numpy()
    c = torch.arange(5, dtype=torch.float32) - 2
    yy, xx = torch.meshgrid(c, c, indexing="ij")
    w["const.bilateral_spatial"] = torch.exp(-(yy ** 2 + xx ** 2) / (2 * 2.0 ** 2)).numpy()
    for t in (4, 8, 16, 32):
        sc = []
        s = 2
        while s <= t:
            sc.append(s)
            s *= 2
        w[f"const.frac_logs_{t}"] = torch.log(torch.tensor(sc, dtype=torch.float32)).numpy()
        w[f"const.frac_w_{t}"] = torch.exp(-0.1 * torch.arange(len(sc), dtype=torch.float32)).numpy()
    w["meta.torch_version"] = np.array(torch.__version__)
    np.savez_compressed(os.path.join(OUT, "weights.npz"), **w)

    for case in CASES:
        o = run_case(*case, morph, ba, qz, A, M, Q)
        hist = np.bincount(o["bit_map_mlp"].astype(int).ravel(), minlength=9)[2:]
        print(f"{case[0]:20s} tile={int(o['cfg'][6])} ht={int(o['cfg'][7])} bits(mlp) hist 2..8 = {hist.tolist()} "
              f"C range [{o['complexity'].min():.3f},{o['complexity'].max():.3f}] "
              f"m range [{o['soft_mask'].min():.3f},{o['soft_mask'].max():.3f}]")
    sz = sum(os.path.getsize(os.path.join(OUT, f)) for f in os.listdir(OUT))
    print("golden bytes:", sz)


if __name__ == "__main__":
    main()
