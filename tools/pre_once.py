"""Developer tool: a few preprocess launches at batch 256 of the bench workload (for ncu)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "image-recognition-pipeline_b200")); sys.path.insert(0, ROOT)
import torch, bench
from irp_b200 import _lib, ops
packed, ids, hw = bench.make_workload(256, seed=0, device=torch.device("cuda:0"))
for _ in range(4):
    x = ops.preprocess(packed.pixels, packed.offsets, packed.hw, packed.max_taps, _lib.LAYOUT_NHWC4P)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); x = ops.preprocess(packed.pixels, packed.offsets, packed.hw, packed.max_taps, _lib.LAYOUT_NHWC4P); e1.record()
torch.cuda.synchronize(); print(f"preprocess 256 images: {e0.elapsed_time(e1)*1e3:.1f} us, max_taps {packed.max_taps}")
