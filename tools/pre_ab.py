"""Developer tool: preprocess time at batch 1024 of the bench workload, with and without the rare big images."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "image-recognition-pipeline_b200")); sys.path.insert(0, ROOT)
import numpy as np, torch, bench
from irp_b200 import _lib, ops
from oracle import synth
dev = torch.device("cuda:0")
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
def run(clip):
    orig = synth.mixed_resolution_sizes
    if clip:
        synth.mixed_resolution_sizes = lambda n, seed=0: np.minimum(orig(n, seed), clip)
    packed, ids, hw = bench.make_workload(1024, seed=0, device=dev)
    synth.mixed_resolution_sizes = orig
    byts = bench.algorithmic_preprocess_bytes(hw)
    fn = lambda: ops.preprocess(packed.pixels, packed.offsets, packed.hw, packed.max_taps, _lib.LAYOUT_NHWC4P)
    for _ in range(3): fn()
    torch.cuda.synchronize(); ms = []
    for _ in range(9):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ms.append(e0.elapsed_time(e1))
    t = float(np.median(ms))
    print(f"clip {clip}: max_taps {packed.max_taps}, {t*1e3:.1f} us per 1024 images, {byts/t/1e6:.0f} GB/s, frac {byts/t/1e6/6550.4:.3f}")
run(0); run(1200); run(500)
