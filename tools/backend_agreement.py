#!/usr/bin/env python
"""Surrogate-vs-cv2 agreement diagnostic on the native analyzer (the counterpart of the reference's
`scripts/backend_agreement.py:47-102`).

For N images it computes, per tile and per metric (fractal / texture / gradient / edge / contour), the descriptors

  * "gpu": the tensorised surrogates, evaluated by the native kernels (`MorphologicalComplexityAnalyzer.
    compute_phi_tiles` of `mcaq_yolo_b200.modules` -> `mcaq_morph_phi` / `mcaq_morph_phi_planes`), and
  * "cv2": the exact OpenCV recipes of the paper's Eq. 21-24 (the reference's `metric_backend='cv2'`,
    `core/morphology.py:110-307, 741-797`), restated here on the HOST with cv2 + numpy -- a diagnostic, not a product
    path: nothing in `mcaq_yolo_b200/` imports this file,

pushes both through the SAME native complexity MLP + bilateral filter, and reports tile-level Pearson r, Spearman rho
and the per-backend means -- the table the reference script prints.  The reference needs `skimage` for the uniform
LBP codes; it is not in this image, so `lbp_uniform_8_1` below restates `local_binary_pattern(P=8, R=1, 'uniform')`
(bilinear samples on the unit circle, >= centre, codes 0..8 for <= 2 transitions, 9 otherwise).

    python tools/backend_agreement.py --synthetic 16 --imgsz 640 [--json out.json]
    python tools/backend_agreement.py --images DIR --n 64 --imgsz 640
    python tools/backend_agreement.py --data data.yaml --split val --n 64      (needs ultralytics, like the reference)
"""
import argparse
import glob
import json
import math
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
METRICS = ["fractal", "texture", "gradient", "edge", "contour"]


# ---- the cv2 recipes (host side) ---------------------------------------------------------------------------------
def tile_size(H: int, grid: int = 8) -> int:
    """Largest power of two <= max(4, H // grid) (morphology.py:359-376)."""
    raw = max(4, H // grid)
    return 1 << (raw.bit_length() - 1)


def canny_otsu(t8):
    """5x5 Gaussian (sigma 1), Otsu's threshold as the high and half of it as the low Canny threshold (Eq. 23)."""
    import cv2
    blurred = cv2.GaussianBlur(t8, (5, 5), 1.0)
    thr, _ = cv2.threshold(blurred, 0, 255, cv2.THRESH_BINARY + cv2.THRESH_OTSU)
    return cv2.Canny(blurred, int(max(0, 0.5 * thr)), int(max(1, thr)))


def fractal_dimension(edge01) -> float:
    """Box counting over scales 2, 4, ..: boxes = area-resized map > 0; weighted (exp(-0.1 k)) least-squares slope of
    log(count + 1) against log(scale), negated and clipped to [1, 2] (Eq. 21 / morphology.py:110-160)."""
    import cv2
    h, w = edge01.shape
    if min(h, w) < 4:
        return 1.0
    pts = []
    for k in range(1, int(math.log2(min(h, w))) + 1):
        s = 2 ** k
        hn, wn = h // s, w // s
        if hn <= 0 or wn <= 0:
            continue
        small = cv2.resize(edge01.astype(np.float32), (wn, hn), interpolation=cv2.INTER_AREA)
        n = int((small > 0).sum())
        if n > 0:
            pts.append((s, n))
    if len(pts) < 2:
        return 1.0
    ls = np.log(np.array([p[0] for p in pts], dtype=np.float64))
    lc = np.log(np.array([p[1] for p in pts], dtype=np.float64) + 1.0)
    wgt = np.exp(-0.1 * np.arange(len(pts)))
    slope = np.polyfit(ls, lc, 1, w=wgt)[0]
    return float(np.clip(-slope, 1.0, 2.0))


def lbp_uniform_8_1(gray) -> np.ndarray:
    """skimage.feature.local_binary_pattern(gray, P=8, R=1, method='uniform'): eight samples on the unit circle
    (bilinear interpolation, zero outside the image), bit = sample >= centre; rotation-invariant uniform code =
    number of set bits when the circular pattern has at most two 0/1 transitions, else P + 1."""
    g = gray.astype(np.float64)
    H, W = g.shape
    pad = np.zeros((H + 2, W + 2), dtype=np.float64)
    pad[1:-1, 1:-1] = g
    bits = []
    for p in range(8):
        ang = 2.0 * math.pi * p / 8.0
        dr, dc = -math.sin(ang), math.cos(ang)
        dr, dc = round(dr, 5), round(dc, 5)
        r0, c0 = math.floor(dr), math.floor(dc)
        fr, fc = dr - r0, dc - c0
        acc = np.zeros((H, W), dtype=np.float64)
        for (rr, cc, wt) in ((r0, c0, (1 - fr) * (1 - fc)), (r0, c0 + 1, (1 - fr) * fc),
                             (r0 + 1, c0, fr * (1 - fc)), (r0 + 1, c0 + 1, fr * fc)):
            if wt == 0.0:
                continue
            rs = np.clip(np.arange(H) + 1 + rr, 0, H + 1)
            cs = np.clip(np.arange(W) + 1 + cc, 0, W + 1)
            acc += wt * pad[np.ix_(rs, cs)]
        bits.append((acc >= g).astype(np.int32))
    bits = np.stack(bits)                                          # (8, H, W)
    trans = np.abs(bits - np.roll(bits, 1, axis=0)).sum(axis=0)
    ones = bits.sum(axis=0)
    return np.where(trans <= 2, ones, 9).astype(np.float64)


def texture_entropy(t8) -> float:
    """Entropy (base 2) of the 10-bin density histogram of the uniform LBP codes, over log2(10) (morphology.py:162-193)."""
    hist, _ = np.histogram(lbp_uniform_8_1(t8).ravel(), bins=10, density=True)
    hist = hist + 1e-10
    p = hist / hist.sum()
    return float(-(p * np.log2(p)).sum() / math.log2(10))


def gradient_variance(t8) -> float:
    """v / (v + 1), v = Var(Gx) + Var(Gy) of the 3x3 Sobel responses on the [0, 1] image (Eq. 22)."""
    import cv2
    g = t8.astype(np.float32)
    if g.max() > 1.5:
        g = g / 255.0
    v = float(np.var(cv2.Sobel(g, cv2.CV_32F, 1, 0, ksize=3)) + np.var(cv2.Sobel(g, cv2.CV_32F, 0, 1, ksize=3)))
    return v / (v + 1.0)


def edge_density(t8) -> float:
    e = canny_otsu(t8)
    return float((e > 0).sum() / e.size)


def contour_complexity(t8) -> float:
    """1 - 1 / max(mean_k P_k^2 / (4 pi A_k), 1) over the external contours (area > 10) of the Gaussian adaptive
    threshold (block 11, C 2) of the tile (Eq. 24 / morphology.py:253-307)."""
    import cv2
    binary = cv2.adaptiveThreshold(t8, 255, cv2.ADAPTIVE_THRESH_GAUSSIAN_C, cv2.THRESH_BINARY, 11, 2)
    contours, _ = cv2.findContours(binary, cv2.RETR_EXTERNAL, cv2.CHAIN_APPROX_SIMPLE)
    ics = []
    for c in contours:
        area = cv2.contourArea(c)
        if area > 10:
            per = cv2.arcLength(c, True)
            if per > 0:
                ics.append(per * per / (4.0 * math.pi * area))
    if not ics:
        return 0.0
    return 1.0 - 1.0 / max(float(np.mean(ics)), 1.0)


def cv2_phi_tiles(images: np.ndarray, grid: int = 8):
    """images (B, C, H, W) float -> phi (B, ht, wt, 8) float32 and the five (B, ht, wt) metrics, per tile on the
    per-image min-max normalised uint8 channel mean (morphology.py:741-797)."""
    B, C, H, W = images.shape
    tile = tile_size(H, grid)
    ht, wt = H // tile, W // tile
    phi = np.zeros((B, ht, wt, 8), dtype=np.float32)
    for b in range(B):
        g = images[b].astype(np.float32).mean(axis=0)
        lo, hi = float(g.min()), float(g.max())
        g8 = ((g - lo) / (hi - lo + 1e-8) * 255.0).astype(np.uint8)
        for i in range(ht):
            for j in range(wt):
                t8 = np.ascontiguousarray(g8[i * tile:(i + 1) * tile, j * tile:(j + 1) * tile])
                p1 = fractal_dimension((canny_otsu(t8) > 0).astype(np.uint8)) / 2.0
                p2, p3, p4, p5 = texture_entropy(t8), gradient_variance(t8), edge_density(t8), contour_complexity(t8)
                phi[b, i, j] = [p1, p2, p3, p4, p5, p1 * p2, p3 * p3, math.sqrt(max(p4 * p5, 0.0))]
    return phi, {m: phi[..., k] for k, m in enumerate(METRICS)}


# ---- statistics ------------------------------------------------------------------------------------------------------
def pearson(a, b) -> float:
    if a.std() < 1e-12 or b.std() < 1e-12:
        return float("nan")
    return float(np.corrcoef(a, b)[0, 1])


def spearman(a, b) -> float:
    ra = np.argsort(np.argsort(a)).astype(np.float64)
    rb = np.argsort(np.argsort(b)).astype(np.float64)
    return pearson(ra, rb)


# ---- images ----------------------------------------------------------------------------------------------------------
def synthetic_images(n: int, size: int, seed: int = 0) -> np.ndarray:
    """Seeded scenes with structure at several scales (smooth shading, discs / boxes / strokes, a textured patch,
    sensor noise): stands in for a dataset, which this sandbox does not have."""
    import cv2
    rng = np.random.default_rng(seed)
    out = np.zeros((n, 3, size, size), dtype=np.float32)
    yy, xx = np.mgrid[0:size, 0:size].astype(np.float32) / size
    for k in range(n):
        img = np.zeros((size, size, 3), dtype=np.float32)
        a, b, c = rng.uniform(-1, 1, 3)
        img += (0.5 + 0.25 * (a * xx + b * yy) + 0.1 * np.sin(6.28 * (c * 3 * xx + yy)))[..., None]
        for _ in range(int(rng.integers(4, 12))):
            col = tuple(float(v) for v in rng.uniform(0, 1, 3))
            kind = int(rng.integers(0, 3))
            p = rng.integers(0, size, 4)
            if kind == 0:
                cv2.circle(img, (int(p[0]), int(p[1])), int(rng.integers(size // 40, size // 6)), col, -1, cv2.LINE_AA)
            elif kind == 1:
                cv2.rectangle(img, (int(p[0]), int(p[1])), (int(p[2]), int(p[3])), col, -1)
            else:
                cv2.line(img, (int(p[0]), int(p[1])), (int(p[2]), int(p[3])), col, int(rng.integers(1, 6)), cv2.LINE_AA)
        x0, y0 = rng.integers(0, size // 2, 2)
        patch = rng.uniform(0, 1, (size // 3, size // 3, 1)).astype(np.float32)
        img[y0:y0 + size // 3, x0:x0 + size // 3] = 0.5 * img[y0:y0 + size // 3, x0:x0 + size // 3] + 0.5 * patch
        img = cv2.GaussianBlur(img, (0, 0), float(rng.uniform(0.3, 1.5)))
        img += rng.normal(0, 0.01, img.shape).astype(np.float32)
        out[k] = np.clip(img, 0, 1).transpose(2, 0, 1)
    return out


def load_images(args) -> np.ndarray:
    import cv2
    if args.synthetic:
        return synthetic_images(args.synthetic, args.imgsz, args.seed)
    if args.images:
        files = sorted(f for ext in ("jpg", "jpeg", "png", "bmp") for f in glob.glob(os.path.join(args.images, "*." + ext)))
        imgs = []
        for f in files[:args.n]:
            im = cv2.imread(f, cv2.IMREAD_COLOR)
            if im is None:
                continue
            im = cv2.resize(im, (args.imgsz, args.imgsz), interpolation=cv2.INTER_LINEAR)
            imgs.append(im[..., ::-1].transpose(2, 0, 1).astype(np.float32) / 255.0)
        if not imgs:
            raise SystemExit("no readable images under %s" % args.images)
        return np.stack(imgs)
    # the reference's own route (needs ultralytics)
    from ultralytics.cfg import get_cfg
    from ultralytics.data import build_yolo_dataset
    from ultralytics.data.utils import check_det_dataset
    data = check_det_dataset(args.data)
    cfg = get_cfg(overrides={"imgsz": args.imgsz, "task": "detect", "mode": "val"})
    ds = build_yolo_dataset(cfg=cfg, img_path=data[args.split], batch=1, data=data, mode="val", rect=False, stride=32)
    imgs = []
    for i in range(min(args.n, len(ds))):
        im = ds[i]["img"].float().numpy()
        imgs.append(im / 255.0 if im.max() > 1.5 else im)
    return np.stack(imgs)


def run(images: np.ndarray, grid: int = 8, batch: int = 16):
    """-> {metric: {pearson, spearman, mean_gpu, mean_cv2}} for the five metrics and the fused complexity map."""
    import torch
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from golden_util import weights
    from mcaq_yolo_b200 import modules as M
    from mcaq_yolo_b200 import constants as K
    from mcaq_yolo_b200 import ops
    analyzer, _, _ = M.build_fixture_modules(weights(), device="cuda", grid_size=grid)
    analyzer.eval()
    cmlp = K.pack_complexity_mlp(analyzer.complexity_mlp)
    per = {m: {"gpu": [], "cv2": []} for m in METRICS}
    fused = {"gpu": [], "cv2": []}
    with torch.no_grad():
        for s in range(0, len(images), batch):
            chunk = images[s:s + batch]
            x = torch.from_numpy(chunk).cuda()
            phi_g, _ = analyzer.compute_phi_tiles(x)                                   # native kernels
            phi_c = torch.from_numpy(cv2_phi_tiles(chunk, grid)[0]).cuda()             # host cv2 recipes
            for name, phi in (("gpu", phi_g), ("cv2", phi_c)):
                for k, m in enumerate(METRICS):
                    per[m][name].append(phi[..., k].reshape(-1).float().cpu().numpy())
                cmap = ops.complexity(phi.float().contiguous(), cmlp, K.device_constants(x.device))   # the SAME MLP
                fused[name].append(cmap.reshape(-1).float().cpu().numpy())
    out = {}
    for m in METRICS + ["fused_C"]:
        src = fused if m == "fused_C" else per[m]
        a, b = np.concatenate(src["gpu"]), np.concatenate(src["cv2"])
        out[m] = {"pearson": pearson(a, b), "spearman": spearman(a, b), "mean_gpu": float(a.mean()), "mean_cv2": float(b.mean())}
    out["_tiles_per_image"] = int(len(np.concatenate(fused["gpu"])) // len(images))
    return out


def main(argv=None):
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("--data", default=None, help="dataset yaml (needs ultralytics)")
    ap.add_argument("--split", default="val", choices=["train", "val", "test"])
    ap.add_argument("--images", default=None, help="directory of images")
    ap.add_argument("--synthetic", type=int, default=0, help="use this many seeded synthetic scenes")
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--n", type=int, default=64)
    ap.add_argument("--imgsz", type=int, default=640)
    ap.add_argument("--grid", type=int, default=8)
    ap.add_argument("--json", default=None)
    args = ap.parse_args(argv)
    if not (args.data or args.images or args.synthetic):
        ap.error("one of --data / --images / --synthetic is required")
    images = load_images(args)
    res = run(images, args.grid)
    print("\nBackend agreement over %d images (native surrogates vs cv2 recipes), %d tiles/image:" % (len(images), res["_tiles_per_image"]))
    print(f"{'metric':<10}{'pearson':>9}{'spearman':>10}{'mean_gpu':>10}{'mean_cv2':>10}")
    for m in METRICS + ["fused_C"]:
        r = res[m]
        print(f"{m:<10}{r['pearson']:>9.3f}{r['spearman']:>10.3f}{r['mean_gpu']:>10.3f}{r['mean_cv2']:>10.3f}")
    if args.json:
        with open(args.json, "w") as f:
            json.dump({"n_images": len(images), "imgsz": args.imgsz, "source": "synthetic" if args.synthetic else (args.images or args.data),
                       "metrics": res}, f, indent=2)
        print("\nWrote", args.json)
    return res


if __name__ == "__main__":
    main()
