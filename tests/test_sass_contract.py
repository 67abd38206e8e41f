"""Static (no GPU) guard of the packed-fp32 arithmetic contract.

ptxas contracts `mul.rn.f32x2` feeding `add.rn.f32x2` into one FFMA2 even under -fmad=false (one
rounding instead of two: the kernels would silently stop matching the reference bit for bit).  The
kernels avoid the pattern (common.cuh: fadd2_sep); this test pins the number of FFMA2 instructions in
the built SASS to the ones the source asks for -- Markstein's two per quantisation, plus the explicit
ffma2 of the d(mask) accumulation -- so a contraction introduced by a source or compiler change is
caught by the CPU suite.  It also checks that the packed instructions are really there."""
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIBDIR = os.path.join(ROOT, "mcaq-yolo_b200", "lib")


@pytest.fixture(scope="module")
def sass():
    import __graft_entry__ as g
    g.build()
    out = {}
    for obj in ("tile_quantize_train.o", "tile_quantize.o"):
        txt = subprocess.run(["cuobjdump", "-sass", os.path.join(LIBDIR, obj)], capture_output=True, text=True).stdout
        for chunk in txt.split("Function : ")[1:]:
            name = chunk.split("\n", 1)[0].strip()
            out[name] = chunk
    return out


def count(sass, pattern, op):
    names = [n for n in sass if re.search(pattern, n)]
    assert len(names) == 1, (pattern, names)
    return len(re.findall(r"\b%s\b" % op, sass[names[0]]))


# (kernel pattern, element pairs per unrolled batch, quantisations per element, extra explicit ffma2 per pair)
CASES = [
    # training forward: 2 quantisations per element, batches of 8 (fp32) / 4 (bf16) channels
    (r"train_fwd_vec_kernelIfLi4ELb1ELb0E", 8 * 2, 2, 0),
    (r"train_fwd_vec_kernelIfLi4ELb0ELb0E", 8 * 2, 2, 0),
    (r"train_fwd_vec_kernelI13__nv_bfloat16Li8ELb1ELb0E", 4 * 4, 2, 0),
    # training backward: batches of 4 channels; with a mask one ffma2 per pair for d(mask)
    (r"train_bwd_vec_kernelIfLi4ELb1ELb0E", 4 * 2, 2, 1),
    (r"train_bwd_vec_kernelIfLi4ELb0ELb0E", 4 * 2, 2, 0),
    (r"train_bwd_vec_kernelI13__nv_bfloat16Li8ELb1ELb0E", 4 * 4, 2, 1),
    # distillation-fused forms: forward batches of 4 (fp32) / 2 (bf16) channels, backward 2
    (r"train_fwd_vec_kernelIfLi4ELb1ELb1E", 4 * 2, 2, 0),
    (r"train_fwd_vec_kernelI13__nv_bfloat16Li8ELb1ELb1E", 2 * 4, 2, 0),
    (r"train_fwd_vec_kernelI13__nv_bfloat16Li8ELb0ELb1E", 2 * 4, 2, 0),
    (r"train_bwd_vec_kernelIfLi4ELb1ELb1E", 2 * 2, 2, 1),
    (r"train_bwd_vec_kernelIfLi4ELb0ELb1E", 2 * 2, 2, 0),
    (r"train_bwd_vec_kernelI13__nv_bfloat16Li8ELb1ELb1E", 2 * 4, 2, 1),
    (r"train_bwd_vec_kernelI13__nv_bfloat16Li8ELb0ELb1E", 2 * 4, 2, 0),
    (r"train_bwd_vec_kernelI13__nv_bfloat16Li8ELb0ELb0E", 4 * 4, 2, 0),
    # inference: unrolled 16-channel walk + generic 8-channel loop, 1 quantisation per element
    (r"tile_quantize_vec_kernelIfLi4ELb1ELb0E", (16 + 8) * 2, 1, 0),
    (r"tile_quantize_vec_kernelI13__nv_bfloat16Li8ELb1ELb0E", (16 + 8) * 4, 1, 0),
    (r"tile_quantize_vec_kernelI13__nv_bfloat16Li8ELb0ELb0E", (16 + 8) * 4, 1, 0),
    (r"tile_quantize_vec_kernelIfLi4ELb0ELb1E", (16 + 8) * 2, 1, 0),
]


@pytest.mark.parametrize("pattern,pairs,quants,extra", CASES)
def test_ffma2_count_is_exactly_the_source(sass, pattern, pairs, quants, extra):
    n = count(sass, pattern, "FFMA2")
    want = pairs * (2 * quants + extra)          # Markstein: r = x - q0*s ; q = q0 + r*rinv
    assert n == want, f"{pattern}: {n} FFMA2 in SASS, {want} in the source (ptxas contraction of a packed mul + add?)"


def test_packed_instructions_present(sass):
    for pattern in (r"tile_quantize_vec_kernelI13__nv_bfloat16Li8ELb1ELb0E", r"train_fwd_vec_kernelIfLi4ELb1ELb0E"):
        assert count(sass, pattern, "FMUL2") > 0 and count(sass, pattern, "FADD2") > 0
