#!/usr/bin/env python
"""Parity on REAL backbone features (random-init minimal YOLOv8n, synthetic images): the native modules against the
reference's modules on CPU (the parity bar) and on CUDA (the reference's own GPU noise), per scale, hooks fed with the
SAME un-quantised feature maps.  python tools/real_feature_parity.py [size] [batch]"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
from harness import ref_model  # noqa: E402

pkg = ref_model.load()
from mcaq_yolo.models.mcaq_yolo import MCAQYOLO  # noqa: E402
from mcaq_yolo_b200 import modules as M  # noqa: E402
from golden_util import weights  # noqa: E402

size = int(sys.argv[1]) if len(sys.argv) > 1 else 640
B = int(sys.argv[2]) if len(sys.argv) > 2 else 2
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
torch.manual_seed(0)
ref = MCAQYOLO("yolov8n", pretrained=False, device="cuda").eval()
W = weights()
sd = lambda d: {k: torch.as_tensor(v) for k, v in d.items()}      # noqa: E731
ref.complexity_analyzer.load_state_dict(sd(W["analyzer"]))
ref.bit_mapper.load_state_dict(sd(W["mapper"]))
feats = {}
hooks = [ref.model.model[i].register_forward_hook(lambda m, i_, o, k=i: feats.__setitem__(k, o.detach())) for i in (4, 6, 9)]
x = torch.rand(B, 3, size, size, device="cuda")
with torch.no_grad():
    ref.model(x)                               # hooks of MCAQ inactive: raw features
for h in hooks:
    h.remove()
a_nat, m_nat, _ = M.build_fixture_modules(W, "cuda")
A_cuda, M_cuda = ref.complexity_analyzer, ref.bit_mapper
import copy
A_cpu, M_cpu = copy.deepcopy(A_cuda).cpu(), copy.deepcopy(M_cuda).cpu()
A_cpu.device = "cpu"
out = {}
with torch.no_grad():
    for k, f in feats.items():
        c_n = a_nat(f); b_n = m_nat(c_n, 1.0)
        c_g = A_cuda(f); b_g = M_cuda(c_g, 1.0)
        c_c = A_cpu(f.cpu()); b_c = M_cpu(c_c, 1.0)
        out[k] = {"shape": list(f.shape), "feat_range": [float(f.min()), float(f.max())],
                  "native_vs_refCPU_bits_equal": float((b_n.cpu() == b_c).float().mean()),
                  "native_vs_refCPU_cpx_maxdiff": float((c_n.cpu() - c_c).abs().max()),
                  "refCUDA_vs_refCPU_bits_equal": float((b_g.cpu() == b_c).float().mean()),
                  "refCUDA_vs_refCPU_cpx_maxdiff": float((c_g.cpu() - c_c).abs().max())}
print(json.dumps(out, indent=1))
