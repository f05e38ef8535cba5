"""Developer tool: one PCA fit + LOF at the BASELINE size (for ncu launch lists)."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "image-recognition-pipeline_b200")); sys.path.insert(0, ROOT)
from irp_b200 import ops
from oracle import synth
n, d, k = 27000, 2048, 50
x = torch.from_numpy(synth.embedding_like(3000, d, seed=1)).cuda().repeat(9, 1).contiguous()
x += 0.3 * torch.randn_like(x)
ids = torch.from_numpy(synth.class_assignment(n, 0)).cuda()
def run():
    shift = x[:256].mean(0).contiguous()
    acc = torch.zeros(1 + d + d * d, dtype=torch.float64, device="cuda")
    ops.cov_accumulate(x, shift, acc[:1], acc[1:1 + d], acc[1 + d:].view(d, d))
    mean, comps, ev = ops.pca_fit(acc[:1], acc[1:1 + d], acc[1 + d:].view(d, d), shift, k)
    z = ops.pca_transform(x, mean, comps)
    a = ops.lof(z, ids, 10, 30, 0.05)
    b = ops.lof(z, None, 1, 75, 0.03)
    torch.cuda.synchronize()
for _ in range(2): run()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); run(); e1.record(); torch.cuda.synchronize()
print(f"pca+lof: {e0.elapsed_time(e1):.2f} ms")
