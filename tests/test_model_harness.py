"""Model-level harness: the UNMODIFIED reference `MCAQYOLO` (models/mcaq_yolo.py:222-589), built on the minimal
random-init YOLOv8 of tests/harness through an `ultralytics` stand-in, driven through
`mcaq_yolo_b200.modules.install()`.

CPU part (where a copy of the reference is reachable: /root/reference or baseline/_ref): backbone discovery,
hook protocol, state_dict round trip.  GPU part: the same model with the reference's own torch modules on CUDA
against the native modules -- bit maps, quantised features, detection-head outputs, calibrate() + freeze."""
import copy

import numpy as np
import pytest
import torch

from harness import ref_model

pkg = ref_model.load()
needs_ref = pytest.mark.skipif(pkg is None, reason="no copy of the reference reachable (/root/reference or baseline/_ref)")


def build(scale="n", device="cpu", **kw):
    from mcaq_yolo.models.mcaq_yolo import MCAQYOLO
    torch.manual_seed(0)
    return MCAQYOLO(f"yolov8{scale}", pretrained=False, device=device, **kw)


@needs_ref
@pytest.mark.parametrize("scale,chans", [("n", [64, 128, 256]), ("s", [128, 256, 512])])
def test_reference_model_builds_on_the_minimal_yolov8(scale, chans):
    m = build(scale)
    assert m.backbone_out_indices == [4, 6, 9]            # _find_backbone_out_indices (351-400)
    assert sorted(m.quantizers.keys()) == ["4", "6", "9"] and len(m._mcaq_hooks) == 3
    m.eval()
    with torch.no_grad():
        out, aux = m(torch.rand(1, 3, 128, 128))
    assert [f.shape[1] for f in aux["quantized_features"]] == chans
    assert aux["feature_layers"] == [4, 6, 9]
    assert isinstance(out, tuple) and len(out[1]) == 3


@needs_ref
def test_install_swaps_modules_and_hooks_keeps_state():
    from mcaq_yolo_b200 import modules as M
    from mcaq_yolo_b200.fused import FusedMcaqHook
    m = build("n")
    # give the lazily created EMA buffers a value, as a trained checkpoint has
    m.train()
    with torch.no_grad():
        m(torch.rand(2, 3, 128, 128))
    sd0 = copy.deepcopy(m.state_dict())
    old_hooks = list(m._mcaq_hooks)
    M.install(m, device="cpu")
    assert isinstance(m.complexity_analyzer, M.MorphologicalComplexityAnalyzer)
    assert isinstance(m.bit_mapper, M.ComplexityToBitMappingNetwork)
    assert all(isinstance(q, M.SpatialAdaptiveQuantization) for q in m.quantizers.values())
    sd1 = m.state_dict()
    assert list(sd0.keys()) == list(sd1.keys()), "state_dict keys changed by install()"
    for k in sd0:
        assert torch.equal(sd0[k], sd1[k]), k
    assert len(m._mcaq_hooks) == 3 and all(h.id != o.id for h, o in zip(m._mcaq_hooks, old_hooks))
    for idx in m.backbone_out_indices:
        hooks = list(m.model.model[idx]._forward_hooks.values())
        assert len(hooks) == 1 and isinstance(hooks[0], FusedMcaqHook) and hooks[0].layer_idx == idx
    # the reference's checkpoint loads back into the swapped model (and vice versa)
    m.load_state_dict(sd0)
    m2 = build("n")
    m2.train()
    with torch.no_grad():
        m2(torch.rand(2, 3, 128, 128))
    m2.load_state_dict(sd1)
    # hooks inactive outside MCAQYOLO.forward: the raw backbone is untouched (no CUDA needed)
    m.eval()
    with torch.no_grad():
        y = m.model(torch.rand(1, 3, 128, 128))
    assert torch.is_tensor(y[0])


@needs_ref
def test_linear_mapping_model_installs():
    from mcaq_yolo_b200 import modules as M
    m = build("n", bit_mapping="linear", normalize_complexity=True)
    M.install(m, device="cpu")
    assert isinstance(m.bit_mapper, M.LinearBitMapper) and m.normalize_complexity is True


# --------------------------------------------------------------------------------------------- GPU
gpu = pytest.mark.gpu


def _pair(scale="n", **kw):
    """(reference-module model, native-module model) with identical weights, on cuda, eval mode."""
    from mcaq_yolo_b200 import modules as M
    torch.backends.cudnn.allow_tf32 = False               # the reference's stencils in fp32, not TF32
    torch.backends.cuda.matmul.allow_tf32 = False
    ref = build(scale, device="cuda", **kw).eval()
    # make the freshly initialised mapper / analyzer spread their outputs a little (all-8 bit maps otherwise)
    from golden_util import weights
    W = weights()
    sd = lambda d: {k: torch.as_tensor(v) for k, v in d.items()}      # noqa: E731
    ref.complexity_analyzer.load_state_dict(sd(W["analyzer"]))
    if "bit_mapping" not in kw:
        ref.bit_mapper.load_state_dict(sd(W["mapper"]))
    for q in ref.quantizers.values():
        q.load_state_dict(sd(W["quantizer"]))
    nat = copy.deepcopy(ref)
    nat._mcaq_hooks = []
    for idx in nat.backbone_out_indices:                 # deepcopy keeps the closures of the ORIGINAL model:
        nat.model.model[idx]._forward_hooks.clear()      # re-register the reference's own hooks on the copy first
    layers = list(nat.model.model)
    nat._mcaq_hooks = [layers[i].register_forward_hook(nat._make_mcaq_hook(i)) for i in nat.backbone_out_indices]
    M.install(nat)
    return ref, nat.eval()


@gpu
@needs_ref
@pytest.mark.parametrize("scale,size,batch", [("n", 640, 2), ("s", 320, 2)])
def test_native_hooks_match_reference_model(scale, size, batch):
    from golden_util import bit_ambiguous  # noqa: F401
    ref, nat = _pair(scale)
    torch.manual_seed(1)
    x = torch.rand(batch, 3, size, size, device="cuda")
    with torch.no_grad():
        o_ref, a_ref = ref(x)
        o_nat, a_nat = nat(x)
    assert a_ref["feature_layers"] == a_nat["feature_layers"] == [4, 6, 9]
    same_bits = True
    for k, (br, bn) in enumerate(zip(a_ref["bit_map"], a_nat["bit_map"])):
        nd = int((br != bn).sum())
        # later scales see features already quantised upstream: identical bit maps upstream => identical inputs
        assert nd <= max(1, br.numel() // 200), f"scale {k}: {nd} of {br.numel()} tiles differ"
        same_bits &= nd == 0
        if same_bits:
            np.testing.assert_allclose(a_nat["complexity_map"][k].cpu().numpy(), a_ref["complexity_map"][k].cpu().numpy(),
                                       rtol=0, atol=2e-2)
    if same_bits:
        for fr, fn in zip(a_ref["quantized_features"], a_nat["quantized_features"]):
            np.testing.assert_allclose(fn.cpu().numpy(), fr.cpu().numpy(), rtol=1e-3, atol=1e-4)
        np.testing.assert_allclose(o_nat[0].cpu().numpy(), o_ref[0].cpu().numpy(), rtol=5e-3, atol=5e-3)
    assert float(a_nat["avg_bits"]) == pytest.approx(float(a_ref["avg_bits"]), abs=0.1)


@gpu
@needs_ref
def test_calibrate_through_the_unmodified_reference_method():
    """MCAQYOLO.calibrate (475-508): eval model, hooks with calibrating=True, EMA then freeze."""
    ref, nat = _pair("n")
    torch.manual_seed(2)
    loader = [(torch.rand(2, 3, 320, 320), None) for _ in range(3)]
    ref.calibrate(loader, num_images=6)
    nat.calibrate(loader, num_images=6)
    for k in ref.quantizers:
        qr, qn = ref.quantizers[k], nat.quantizers[k]
        assert bool(qr.stats_frozen) and bool(qn.stats_frozen) and int(qn.num_batches_tracked) == 3
        if k == "4":       # C3 statistics: the hooked layer's input is identical in both models (no hook upstream)
            np.testing.assert_allclose(qn.running_min.cpu().numpy(), qr.running_min.cpu().numpy(), rtol=1e-4, atol=1e-5)
            np.testing.assert_allclose(qn.running_max.cpu().numpy(), qr.running_max.cpu().numpy(), rtol=1e-4, atol=1e-5)
        else:              # C4 / C5 see C3 features quantised with possibly different bit maps upstream
            np.testing.assert_allclose(qn.running_min.cpu().numpy(), qr.running_min.cpu().numpy(), rtol=0.2, atol=0.2)
    x = torch.rand(2, 3, 320, 320, device="cuda")
    with torch.no_grad():
        _, a_ref = ref(x)
        _, a_nat = nat(x)
    if all(torch.equal(a, b) for a, b in zip(a_ref["bit_map"], a_nat["bit_map"])):
        for fr, fn in zip(a_ref["quantized_features"], a_nat["quantized_features"]):
            np.testing.assert_allclose(fn.cpu().numpy(), fr.cpu().numpy(), rtol=1e-3, atol=1e-4)


@gpu
@needs_ref
def test_model_under_autocast_fp16():
    """The reference trainer runs eval under autocast (train.py:748): hooked outputs are fp16."""
    _, nat = _pair("n")
    x = torch.rand(2, 3, 320, 320, device="cuda")
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.float16):
        out, aux = nat(x)
    assert all(f.dtype == torch.float16 for f in aux["quantized_features"])
    assert all(torch.isfinite(f.float()).all() for f in aux["quantized_features"])
