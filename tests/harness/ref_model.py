"""Import the UNMODIFIED reference package with the offline stubs (skimage.feature, ultralytics).

Search order: $MCAQ_REF, /root/reference (build container), <repo>/baseline/_ref (the reference pip-installed
with `--target baseline/_ref`, git-ignored, travels to the GPU box with the snapshot).  Test / bench
infrastructure only: nothing under mcaq_yolo_b200/ imports this."""
import os
import sys
import tempfile
import warnings

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def reference_root():
    for cand in (os.environ.get("MCAQ_REF"), "/root/reference", os.path.join(ROOT, "baseline", "_ref")):
        if cand and os.path.isdir(os.path.join(cand, "mcaq_yolo", "core")):
            return cand
    return None


def load(with_model=True):
    """Returns the reference's `mcaq_yolo` package (core eager; `models.mcaq_yolo` imported when with_model)
    or None when no copy of the reference is reachable."""
    root = reference_root()
    if root is None:
        return None
    stub = os.path.join(tempfile.gettempdir(), "mcaq_ref_stubs")
    os.makedirs(os.path.join(stub, "skimage"), exist_ok=True)
    open(os.path.join(stub, "skimage", "__init__.py"), "a").close()
    with open(os.path.join(stub, "skimage", "feature.py"), "w") as f:
        f.write("def local_binary_pattern(*a, **k):\n    raise NotImplementedError\n")   # cv2 backend only
    for p in (stub, root):
        if p not in sys.path:
            sys.path.insert(0, p)
    from . import ultralytics_stub
    ultralytics_stub.install()
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        import mcaq_yolo
        if with_model:
            import mcaq_yolo.models.mcaq_yolo  # noqa: F401
    return mcaq_yolo
