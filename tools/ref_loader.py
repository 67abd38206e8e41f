"""Import the real reference (read-only, /root/reference or $MCAQ_REF) in the build
container.  Used ONLY by tools/make_golden.py and by the container-only tests that
cross-check state_dict compatibility.  Never imported by the product or on the GPU box."""
import os
import sys
import tempfile
import warnings


def reference_root():
    for cand in (os.environ.get("MCAQ_REF"), "/root/reference"):
        if cand and os.path.isdir(os.path.join(cand, "mcaq_yolo", "core")):
            return cand
    return None


def load_reference():
    """Returns the `mcaq_yolo.core` submodules (morphology, bit_allocation, quantization)
    or raises RuntimeError when the reference tree is absent."""
    root = reference_root()
    if root is None:
        raise RuntimeError("reference tree not found (set MCAQ_REF or mount /root/reference)")
    stub = os.path.join(tempfile.gettempdir(), "mcaq_ref_stubs")
    os.makedirs(os.path.join(stub, "skimage"), exist_ok=True)
    open(os.path.join(stub, "skimage", "__init__.py"), "a").close()
    with open(os.path.join(stub, "skimage", "feature.py"), "w") as f:
        # only the out-of-scope cv2 backend calls this (morphology.py:13, 181)
        f.write("def local_binary_pattern(*a, **k):\n    raise NotImplementedError\n")
    for p in (stub, root):
        if p not in sys.path:
            sys.path.insert(0, p)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        from mcaq_yolo.core import morphology, bit_allocation, quantization
    return morphology, bit_allocation, quantization
