"""mcaq_yolo_b200: B200-native (sm_100a) implementation of MCAQ-YOLO's data-parallel hot path.

Host side mirrors the reference's interface for that path (same class names, constructor
arguments, call signatures and state_dict keys as mcaq_yolo/core/*.py) on top of the C-ABI
library libmcaq_b200.so (include/mcaq_b200.h).  There is no CPU fallback: every op raises
if the CUDA library or a CUDA device is missing.
"""
__version__ = "0.1.0"
