"""Developer tool: whole-trunk time per 256 images as a function of the per-call batch."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "image-recognition-pipeline_b200")); sys.path.insert(0, ROOT)
from irp_b200.stage import ResNet50Trunk
from oracle import stage_ref
model = stage_ref.full_resnet50(1234)
xp = torch.zeros(256, 230, 230, 4, device="cuda", dtype=torch.bfloat16)
xp[:, 3:227, 3:227, :3] = torch.randn(256, 224, 224, 3, device="cuda").bfloat16()
for B in [int(a) for a in sys.argv[1:]] or [256, 128, 64, 32, 16]:
    trunk = ResNet50Trunk(model, torch.device("cuda:0"), max_batch=B)
    parts = [xp[s:s + B].contiguous() for s in range(0, 256, B)]
    def run():
        for p in parts: trunk.embed(p)
    for _ in range(3): run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): run()
    e1.record(); torch.cuda.synchronize()
    print(f"batch {B:4d}: {e0.elapsed_time(e1) / 5:.3f} ms per 256 images", flush=True)
    trunk.close()
