"""The quantiser's fast division (Markstein correction with RN(1/scale)) must equal div.rn
bit-for-bit: swept on the GPU over every fp32 numerator pattern (stride 1 over a window, strided
over the whole range) for random, adversarial (all-ones / power-of-two significand) and
fixture-derived scales."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _run(scales, first, stride, count):
    from mcaq_yolo_b200 import _lib
    lib = _lib.load()
    s = torch.tensor(scales, dtype=torch.float32, device="cuda")
    bad = torch.zeros(2, dtype=torch.int64, device="cuda")
    rc = lib.mcaq_selftest_division(s.data_ptr(), s.numel(), first, stride, count, bad.data_ptr(),
                                    torch.cuda.current_stream().cuda_stream)
    assert rc == 0
    torch.cuda.synchronize()
    n, ex = int(bad[0].item()), int(bad[1].item())
    if n:
        xb, sb = (ex >> 32) & 0xFFFFFFFF, ex & 0xFFFFFFFF
        x = np.array([xb], dtype=np.uint32).view(np.float32)[0]
        sc = np.array([sb], dtype=np.uint32).view(np.float32)[0]
        print(f"division mismatch example: x={x!r} ({xb:#x}) scale={sc!r} ({sb:#x}), {n} mismatches")
    return n


def _scales():
    rng = np.random.Generator(np.random.PCG64(11))
    rnd = np.exp(rng.uniform(np.log(1e-6), np.log(1e3), 96)).astype(np.float32)
    adv = []
    for e in (-20, -8, -3, 0, 1, 5):
        for mant in (0x7FFFFF, 0x7FFFFE, 0x000000, 0x000001, 0x400000, 0x3FFFFF, 0x555555, 0x2AAAAA):
            bits = ((127 + e) << 23) | mant
            adv.append(np.array([bits], dtype=np.uint32).view(np.float32)[0])
    rng_scales = [(mx - mn) / (2 ** b - 1) for b in range(2, 9) for mn, mx in ((-3.7, 5.1), (0.0, 1.0), (-1e-3, 2e-3))]
    return [float(v) for v in rnd] + [float(v) for v in adv] + [float(np.float32(v)) for v in rng_scales]


def test_markstein_division_matches_div_rn_strided_full_range():
    # every 1021st bit pattern over all 2^32 numerators, all scales
    assert _run(_scales(), 0, 1021, (1 << 32) // 1021) == 0


def test_markstein_division_matches_div_rn_dense_windows():
    scales = _scales()
    # dense (stride 1) windows of 2^24 patterns around typical activation magnitudes, both signs
    for first in (0x3F000000, 0x40400000, 0xBF800000, 0xC1200000, 0x3A000000, 0x00000000, 0x00800000):
        assert _run(scales, first, 1, 1 << 24) == 0, hex(first)
