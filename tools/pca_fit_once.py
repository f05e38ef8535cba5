"""Developer tool: one PCA fit at d=2048, k=50 (for ncu launch lists)."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "image-recognition-pipeline_b200")); sys.path.insert(0, ROOT)
from irp_b200 import ops
from oracle import synth
d, k = 2048, 50
x = torch.from_numpy(synth.embedding_like(3000, d, seed=1)).cuda()
shift = x[:256].mean(0).contiguous()
acc = torch.zeros(1 + d + d * d, dtype=torch.float64, device="cuda")
ops.cov_accumulate(x, shift, acc[:1], acc[1:1 + d], acc[1 + d:].view(d, d))
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 1
for _ in range(reps):
    mean, comps, ev = ops.pca_fit(acc[:1], acc[1:1 + d], acc[1 + d:].view(d, d), shift, k)
torch.cuda.synchronize()
import time
t = time.perf_counter()
for _ in range(reps):
    mean, comps, ev = ops.pca_fit(acc[:1], acc[1:1 + d], acc[1 + d:].view(d, d), shift, k)
torch.cuda.synchronize()
print(f"wall per fit {(time.perf_counter() - t) / reps * 1e3:.2f} ms")
