"""Developer tool: a few whole-trunk calls at batch 256 (for ncu launch lists / captures)."""
import ctypes as C, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "image-recognition-pipeline_b200")); sys.path.insert(0, ROOT)
from irp_b200 import _lib, ops
if os.environ.get("IRP_AB_LIB"):  # developer A/B against another build of the library (tools/ab_trunk.sh)
    _lib.LIB_PATH = os.path.join(ROOT, os.environ["IRP_AB_LIB"])
from irp_b200.stage import ResNet50Trunk
from oracle import stage_ref
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 3
trunk = ResNet50Trunk(stage_ref.full_resnet50(1234), torch.device("cuda:0"), max_batch=B)
xp = torch.zeros(B, 230, 230, 4, device="cuda", dtype=torch.bfloat16)
xp[:, 3:227, 3:227, :3] = torch.randn(B, 224, 224, 3, device="cuda").bfloat16()
for _ in range(iters):
    out = trunk.embed(xp)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(iters):
    out = trunk.embed(xp)
e1.record(); torch.cuda.synchronize()
print(f"batch {B}: {e0.elapsed_time(e1)/iters:.3f} ms per trunk call")
