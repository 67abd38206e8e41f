import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
from mcaq_yolo_b200 import modules as M
from golden_util import weights
from harness import ref_model
ref_model.load(with_model=False)
from mcaq_yolo.core import morphology
W = weights()
a, _, _ = M.build_fixture_modules(W, "cuda")
ref_a = morphology.MorphologicalComplexityAnalyzer(grid_size=8, device="cuda")
ref_a.load_state_dict({k: torch.as_tensor(v) for k, v in W["analyzer"].items()})
ref_a = ref_a.to("cuda").eval()
ref_c = morphology.MorphologicalComplexityAnalyzer(grid_size=8, device="cpu")
ref_c.load_state_dict({k: torch.as_tensor(v) for k, v in W["analyzer"].items()})
ref_c.eval()
torch.manual_seed(0)
x = torch.rand(6, 3, 640, 640, device="cuda")
with torch.no_grad():
    sb = a.score_image(x)
    ss = torch.cat([a.score_image(x[i:i + 1]) for i in range(6)])
    print("native batch vs native single:", (sb - ss).abs().max().item())
    for tf32 in (True, False):
        torch.backends.cudnn.allow_tf32 = tf32
        torch.backends.cuda.matmul.allow_tf32 = tf32
        sr = torch.cat([ref_a.score_image(x[i:i + 1]) for i in range(6)])
        print("tf32", tf32, "native vs reference CUDA:", (sb - sr).abs().tolist())
    sc = torch.cat([ref_c.score_image(x[i:i + 1].cpu()) for i in range(6)])
    print("native vs reference CPU (6 images):", (sb.cpu() - sc).abs().tolist())
