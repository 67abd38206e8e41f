"""Stand-in for the three names the reference imports from `ultralytics` (models/mcaq_yolo.py:9-11):
`YOLO`, `utils.loss.v8DetectionLoss`, `cfg.DEFAULT_CFG`.  `ultralytics==8.4.63` is pinned by the reference but is
neither vendored nor installable here (no network).  `install()` registers the stub in sys.modules only when the
real package is absent."""
import re
import sys
import types
from types import SimpleNamespace

from .mini_yolov8 import DetectionModel


class YOLO:
    """`YOLO('yolov8n')` / `YOLO('yolov8s.pt')`: a wrapper whose `.model` is a random-init DetectionModel (there
    are no checkpoints offline, so a `.pt` name also yields random weights -- said in every result's `data`)."""

    def __init__(self, name="yolov8n", task=None):
        m = re.match(r"yolov8([nsm])", str(name))
        if not m:
            raise ValueError(f"ultralytics stub: only yolov8n/s/m are available offline, got {name!r}")
        self.model = DetectionModel(scale=m.group(1))
        self.task = "detect"


class v8DetectionLoss:
    """Constructible (MCAQYOLOLoss.__init__ builds one, models/mcaq_yolo.py:85); calling it needs the real
    assigner / DFL machinery, which is out of scope: Ldet is omitted and said so (SURVEY 8d configs[3])."""

    def __init__(self, model, *a, **k):
        self.model = model

    def __call__(self, preds, batch):
        raise NotImplementedError("v8DetectionLoss is not available offline (ultralytics stub)")


DEFAULT_CFG = SimpleNamespace(box=7.5, cls=0.5, dfl=1.5, pose=12.0, kobj=1.0)


def install():
    try:
        import ultralytics  # noqa: F401  (the real package wins when present)
        return False
    except ImportError:
        pass
    root = types.ModuleType("ultralytics")
    root.YOLO = YOLO
    root.__stub__ = True
    utils = types.ModuleType("ultralytics.utils")
    loss = types.ModuleType("ultralytics.utils.loss")
    loss.v8DetectionLoss = v8DetectionLoss
    cfg = types.ModuleType("ultralytics.cfg")
    cfg.DEFAULT_CFG = DEFAULT_CFG
    utils.loss = loss
    root.utils, root.cfg = utils, cfg
    sys.modules.update({"ultralytics": root, "ultralytics.utils": utils, "ultralytics.utils.loss": loss,
                        "ultralytics.cfg": cfg})
    return True
