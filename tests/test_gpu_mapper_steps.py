"""Step table of the MLP bit mapper (mcaq_mapper_steps): the staircase read by the fused kernel must
reproduce the mapper kernel's own integer outputs (which tests/test_gpu_parity.py pins bit-exactly to
the reference) for every complexity value, including the two fp32 neighbours of every step."""
import numpy as np
import pytest
import torch

from golden_util import weights

pytestmark = pytest.mark.gpu


def _mods():
    from mcaq_yolo_b200 import modules as M
    return M.build_fixture_modules(weights(), "cuda")


@pytest.mark.parametrize("temperature", [1.0, 0.7, 1.3, None])
def test_step_table_equals_mapper_kernel(temperature):
    from mcaq_yolo_b200 import constants as K
    from mcaq_yolo_b200 import ops
    a, m, q = _mods()
    ext = K.pack_mapping_steps(m.mapping_network, temperature, m.min_bits, m.max_bits)
    assert K.pack_mapping_steps(m.mapping_network, temperature, m.min_bits, m.max_bits) is ext      # cached
    tab = ext[ops.MAPPER_FLOATS:].cpu().numpy()
    steps = tab[:8]
    assert tab[8] == 1.0, "fixture mapper is monotone: table must be valid"
    assert tab[10] == m.min_bits and tab[11] == m.max_bits
    assert np.all(steps[1:] >= steps[:-1])
    finite = steps[np.isfinite(steps) & (steps > 0)]
    assert finite.size >= 3, "spread-mapper fixture crosses several bit widths"
    # dense sample + the fp32 neighbourhood of every step
    rng = np.random.default_rng(5)
    c = [rng.random(200_000, dtype=np.float32), np.linspace(0, 1, 4097, dtype=np.float32)]
    for s in finite:
        u = np.float32(s).view(np.uint32)
        c.append(np.arange(int(u) - 64, int(u) + 65, dtype=np.int64).astype(np.uint32).view(np.float32))
    c = np.clip(np.concatenate(c), 0, 1).astype(np.float32)
    n = (c.size // 64) * 64
    c = c[:n]
    ref = ops.bit_mapper(torch.from_numpy(c).cuda().reshape(n // 64, 8, 8), K.pack_mapping_network(m.mapping_network),
                         temperature, False, m.min_bits, m.max_bits).cpu().numpy().reshape(-1)
    stair = m.min_bits + (c[:, None] >= steps[None, :]).sum(1)
    bad = np.flatnonzero(ref != stair)
    # by construction the two agree AT each step and just below it; a non-monotone wiggle of the network
    # within a few ulps of a step is the only place they may differ (continuous bit value on a .5 boundary)
    for i in bad:
        d = np.abs(c[i].view(np.uint32).astype(np.int64) - finite.view(np.uint32).astype(np.int64)).min()
        assert d <= 64, f"staircase differs from the mapper away from a step: c={c[i]!r}"
    assert bad.size <= 8, f"{bad.size} disagreements near steps"
    for s in finite:                                    # exact at the step and at its predecessor
        pair = np.array([np.nextafter(np.float32(s), np.float32(0)), s] * 4, dtype=np.float32)
        r = ops.bit_mapper(torch.from_numpy(pair).cuda().reshape(1, 1, 8), K.pack_mapping_network(m.mapping_network),
                           temperature, False, m.min_bits, m.max_bits).cpu().numpy().reshape(-1)
        assert r[1] == r[0] + 1 or r[1] > r[0]


def test_table_is_rebuilt_when_weights_change():
    from mcaq_yolo_b200 import constants as K
    a, m, q = _mods()
    e1 = K.pack_mapping_steps(m.mapping_network, 1.0, 2.0, 8.0)
    t1 = e1.clone()
    with torch.no_grad():
        m.mapping_network[9].bias.add_(0.75)
    e2 = K.pack_mapping_steps(m.mapping_network, 1.0, 2.0, 8.0)
    assert e2 is not e1 and not torch.equal(e2[-12:-4], t1[-12:-4])


def test_fused_kernel_uses_table_and_matches_network(monkeypatch):
    """Same launch with the plain block (network evaluated per tile) and with the table: identical maps."""
    from mcaq_yolo_b200 import constants as K
    from mcaq_yolo_b200 import ops
    from golden_util import Case
    a, m, q = _mods()
    for name in ("c3_v8n_smooth", "c4_v8n_smooth", "c5_v8n_smooth", "c3_v8n_noise"):
        c = Case(name)
        x = torch.from_numpy(c.x()).cuda()
        s, ab, keys = ops.reduce_planes(x)
        cm, sm = K.pack_complexity_mlp(a.complexity_mlp), K.pack_soft_mask(q.soft_mask)
        r0 = ops.morph_fused(s, ab, c.C, c.grid, cm, K.pack_mapping_network(m.mapping_network), sm, 1.0)
        r1 = ops.morph_fused(s, ab, c.C, c.grid, cm, K.pack_mapping_steps(m.mapping_network, 1.0, 2.0, 8.0), sm, 1.0)
        assert torch.equal(r0["bit_map"], r1["bit_map"]) and torch.equal(r0["mask"], r1["mask"])
        assert np.array_equal(r1["bit_map"].cpu().numpy(), c["bit_map_mlp"]), "reference bit map"
        # continuous output ignores the table
        rc0 = ops.morph_fused(s, ab, c.C, c.grid, cm, K.pack_mapping_network(m.mapping_network), sm, 1.3, True)
        rc1 = ops.morph_fused(s, ab, c.C, c.grid, cm, K.pack_mapping_steps(m.mapping_network, 1.3, 2.0, 8.0), sm, 1.3, True)
        assert torch.equal(rc0["bit_map"], rc1["bit_map"])
