"""Model-level test / bench harness (NOT product code): a minimal random-init YOLOv8 and an `ultralytics`
stand-in, so that the UNMODIFIED reference `MCAQYOLO` (models/mcaq_yolo.py:222-589) can be constructed
without the un-vendored `ultralytics` dependency and driven through `mcaq_yolo_b200.modules.install()`."""
