#!/usr/bin/env python
"""Round-2 additions to the committed reference outputs (tests/golden/r2/*.npz), produced by RUNNING
THE REAL REFERENCE (imported read-only from /root/reference) like tools/make_golden.py:

  calib_c4      MCAQYOLO.calibrate()'s use of the quantizer: eval-mode module, training=True
                (models/mcaq_yolo.py:446, quantization.py:415-417): two calibration batches, freeze, one
                inference batch -- outputs and running statistics after every call
  normalize_c3  the hook with normalize_complexity=True (models/mcaq_yolo.py:427-432)
  nonmono_c3    an MLP mapper that is NOT monotone (enforce_monotonicity=False, signed weights)
  image_640 / image_1280 / image_320
                compute_phi_tiles / score_image on image-sized inputs (utils/dataset.py:345-353,
                tests/test_smoke.py:33-47): tile 64 / 128 / 32

Run in the build container only:  python tools/make_golden_r2.py
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
from inputs import feature_map, integer_bit_map  # noqa: E402
from golden_util import weights  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden", "r2")


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def sd(d):
    return {k: torch.as_tensor(v) for k, v in d.items()}


def fixtures(morph, ba, qz, grid=8):
    W = weights()
    A = morph.MorphologicalComplexityAnalyzer(grid_size=grid, device="cpu")
    A.load_state_dict(sd(W["analyzer"]))
    M = ba.ComplexityToBitMappingNetwork()
    M.load_state_dict(sd(W["mapper"]))
    Q = qz.SpatialAdaptiveQuantization(calibration_mode="minmax", smooth_transitions=True, per_channel=True)
    Q.load_state_dict(sd(W["quantizer"]))
    return A.eval(), M.eval(), Q.eval()


def calib_case(morph, ba, qz):
    A, M, Q = fixtures(morph, ba, qz)
    B, C, H = 2, 128, 40
    xs = [torch.from_numpy(feature_map("smooth", B, C, H, H, 31 + i)) * (1.0 + 0.5 * i) + 0.25 * i for i in range(3)]
    bm = torch.from_numpy(integer_bit_map(B, 10, 10, 31))
    out = dict(cfg=np.array([B, C, H, H, 8, 31], dtype=np.int64), bit_map=bm.numpy())
    Q.eval()                                   # MCAQYOLO.calibrate(): self.eval(), hooks pass training=True
    with torch.no_grad():
        for i in range(2):
            y = Q(xs[i], bm, training=True)
            out[f"y{i}_sha"] = np.array(sha(y.numpy()))
            out[f"y{i}_sub"] = y.numpy()[:, ::3, ::5, ::7]
            out[f"run_min{i}"] = Q.running_min.numpy().ravel().copy()
            out[f"run_max{i}"] = Q.running_max.numpy().ravel().copy()
        Q.freeze_calibration()
        y = Q(xs[2], bm, training=False)
        out["y2_sha"] = np.array(sha(y.numpy()))
        out["y2_sub"] = y.numpy()[:, ::3, ::5, ::7]
    np.savez_compressed(os.path.join(OUT, "calib_c4.npz"), **out)
    print("calib_c4", out["y0_sha"], out["y1_sha"])


def normalize_case(morph, ba, qz):
    A, M, Q = fixtures(morph, ba, qz)
    B, C, H = 2, 64, 80
    x = torch.from_numpy(feature_map("smooth", B, C, H, H, 41))
    with torch.no_grad():
        complexity = A(x)
        # the hook's optional normalisation, verbatim semantics of models/mcaq_yolo.py:427-432
        flat = complexity.reshape(B, -1)
        lo = torch.quantile(flat, 0.02, dim=1, keepdim=True).unsqueeze(-1)
        hi = torch.quantile(flat, 0.98, dim=1, keepdim=True).unsqueeze(-1)
        cn = ((complexity - lo) / (hi - lo + 1e-8)).clamp(0.0, 1.0)
        bm = M(cn, 1.0)
        bl = ba.LinearBitMapper()(cn, 1.0)
        Q.eval()
        y = Q(x, bm, training=False)
    np.savez_compressed(os.path.join(OUT, "normalize_c3.npz"), cfg=np.array([B, C, H, H, 8, 41], dtype=np.int64),
                        complexity=complexity.numpy(), complexity_norm=cn.numpy(), bit_map_mlp=bm.numpy(),
                        bit_map_linear=bl.numpy(), y_sha=np.array(sha(y.numpy())), y_sub=y.numpy()[:, ::3, ::5, ::7])
    print("normalize_c3 bits", np.bincount(bm.numpy().astype(int).ravel(), minlength=9)[2:])


def nonmono_case(morph, ba, qz):
    A, M0, Q = fixtures(morph, ba, qz)
    torch.manual_seed(77)
    M = ba.ComplexityToBitMappingNetwork(enforce_monotonicity=False)
    M.load_state_dict(M0.state_dict())
    M.eval()
    with torch.no_grad():
        # first layer: negative weight on c^2, none on log1p(c) -> every unit rises then falls on [0, 1];
        # last layer re-centred on a complexity ramp so that the bump spans several bit widths
        ramp = torch.linspace(0, 1, 400).reshape(-1, 1)
        w0 = M.mapping_network[0].weight
        w0[:, 0] = w0[:, 0].abs() * 4.0 + 0.5
        w0[:, 1] = -1.1 * w0[:, 0]
        w0[:, 2] = 0.0
        z = M.mapping_network[:-1](M.create_augmented_features(ramp))
        sc = 2.5 / z.std()
        b9 = M.mapping_network[9].bias.clone()
        M.mapping_network[9].weight.mul_(sc)
        M.mapping_network[9].bias.copy_(sc * b9 - sc * z.mean())
    M.eval()
    c = torch.linspace(0, 1, 400).reshape(1, 20, 20)
    with torch.no_grad():
        b = M(c, 1.0)
        bc = M(c, 1.0, return_continuous=True)
    d = np.diff(b.numpy().ravel())
    assert (d > 0).any() and (d < 0).any(), "fixture is monotone after all"
    w = {("mapper." + k): v.numpy() for k, v in M.state_dict().items()}
    x = torch.from_numpy(feature_map("smooth", 1, 64, 80, 80, 43))
    with torch.no_grad():
        cpx = A(x)
        bx = M(cpx, 1.0)
        bxc = M(cpx, 1.0, return_continuous=True)
    np.savez_compressed(os.path.join(OUT, "nonmono_c3.npz"), ramp_bits=b.numpy(), ramp_bits_cont=bc.numpy(),
                        cfg=np.array([1, 64, 80, 80, 8, 43], dtype=np.int64), bit_map=bx.numpy(),
                        bit_map_cont=bxc.numpy(), complexity=cpx.numpy(), **w)
    print("nonmono_c3 ramp bits", np.bincount(b.numpy().astype(int).ravel(), minlength=9)[2:],
          "map bits", np.bincount(bx.numpy().astype(int).ravel(), minlength=9)[2:])


def image_case(morph, ba, qz, name, S, seed, kind="smooth", planes=True):
    A, _, _ = fixtures(morph, ba, qz)
    x = torch.from_numpy(feature_map(kind, 1, 3, S, S, seed))
    tile = A._tile_size(S)
    ht = S // tile
    Hc = ht * tile
    with torch.no_grad():
        gray = A._normalize01(x[:, :, :Hc, :Hc].mean(dim=1, keepdim=True).float())
        edge = A._gpu_canny(gray)
        binm = A._binarize(gray)
        phi, _ = A.compute_phi_tiles(x)
        score = A.score_image(x)
        cpx = A(x)
    out = dict(cfg=np.array([1, 3, S, S, 8, seed, tile, ht, ht], dtype=np.int64), kind=np.array(kind),
               phi=phi.numpy(), score=score.numpy(), complexity=cpx.numpy(),
               edge_sha=np.array(sha(np.packbits(edge[:, 0].numpy() > 0))),
               bin_sha=np.array(sha(np.packbits(binm[:, 0].numpy() > 0))),
               edge_count=np.array(int(edge.sum())), bin_count=np.array(int(binm.sum())),
               gray_sub=gray[0, 0, ::7, ::5].numpy())
    if planes:
        out["edge"] = np.packbits(edge[:, 0].numpy() > 0)
        out["binmask"] = np.packbits(binm[:, 0].numpy() > 0)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
    print(name, "tile", tile, "ht", ht, "edges", int(edge.sum()), "bin", int(binm.sum()),
          "phi range", [float(phi[..., i].min()) for i in range(5)], [float(phi[..., i].max()) for i in range(5)])


def main():
    morph, ba, qz = load_reference()
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(8)
    which = sys.argv[1:] or ["calib", "normalize", "nonmono", "images"]
    if "calib" in which:
        calib_case(morph, ba, qz)
    if "normalize" in which:
        normalize_case(morph, ba, qz)
    if "nonmono" in which:
        nonmono_case(morph, ba, qz)
    if "images" in which:
        image_case(morph, ba, qz, "image_320", 320, 51)
        image_case(morph, ba, qz, "image_640", 640, 52)
        image_case(morph, ba, qz, "image_640_noise", 640, 53, kind="noise", planes=False)
        image_case(morph, ba, qz, "image_1280", 1280, 54, planes=False)
    print("bytes:", sum(os.path.getsize(os.path.join(OUT, f)) for f in os.listdir(OUT)))


if __name__ == "__main__":
    main()
