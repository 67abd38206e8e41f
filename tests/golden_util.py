"""Load the committed reference outputs (tests/golden/*.npz, written by tools/make_golden.py)."""
import hashlib
import os

import numpy as np

from inputs import feature_map

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

CASE_NAMES = sorted(f[:-4] for f in os.listdir(GOLDEN_DIR) if f.endswith(".npz") and f != "weights.npz")


def sha(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


_W = None


def weights():
    """dict(analyzer=..., mapper=..., quantizer=..., const=...) of fp32 arrays keyed by state_dict name."""
    global _W
    if _W is None:
        z = np.load(os.path.join(GOLDEN_DIR, "weights.npz"))
        out = {"analyzer": {}, "mapper": {}, "quantizer": {}, "const": {}, "meta": {}}
        for k in z.files:
            grp, name = k.split(".", 1)
            out[grp][name] = z[k]
        _W = out
    return _W


class Case:
    def __init__(self, name):
        self.name = name
        self.z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
        (self.B, self.C, self.H, self.W, self.grid, self.seed,
         self.tile, self.ht, self.wt) = [int(v) for v in self.z["cfg"]]
        self.kind = str(self.z["kind"])
        self.Hc, self.Wc = self.ht * self.tile, self.wt * self.tile

    def x(self):
        return feature_map(self.kind, self.B, self.C, self.H, self.W, self.seed)

    def grad(self):
        return feature_map("noise", self.B, self.C, self.H, self.W, self.seed + 500)

    def plane_bits(self, key):
        n = self.B * self.Hc * self.Wc
        return np.unpackbits(self.z[key])[:n].reshape(self.B, self.Hc, self.Wc).astype(bool)

    def __getitem__(self, k):
        return self.z[k]


def bit_ambiguous(pre_round: np.ndarray, tol: float = 1e-4) -> np.ndarray:
    """Tiles whose continuous bit value sits within `tol` of a rounding boundary: the
    integer bit there is not well defined across ulp-level summation-order changes
    (SURVEY 7.3 'oracle-ambiguity set')."""
    frac = pre_round - np.floor(pre_round)
    return np.abs(frac - 0.5) < tol
