#!/usr/bin/env python
"""Print the mapper step table K2 uses for the fixture weights (valid flag at [8]): python tools/steps_probe.py"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from golden_util import weights
from mcaq_yolo_b200 import modules as M, constants as K
from mcaq_yolo_b200.fused import mapper_block
a, m, q = M.build_fixture_modules(weights(), device="cuda")
for t in (1.0, None, 0.5):
    blk = mapper_block(m, t)
    print("temperature", t, "monotone", K.mapping_is_monotone(m), "numel", blk.numel(), "steps", blk[-12:].tolist())
