"""Measurement for SURVEY.md 8f row N1 (classifier inference): images/s of val_transform + ResNet-50 + head +
cross-entropy statistics over the same 27 000 mixed-resolution synthetic images as bench.py, device-resident and
end to end from pinned host memory, next to the reference's CPU route (oracle/classifier_ref.py) on a sample.

    python tools/bench_classify.py [--steps K] [--warmup W]      -> one JSON line
"""
import argparse, json, os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "image-recognition-pipeline_b200")); sys.path.insert(0, ROOT)
import bench  # workload generator, clocks sampler, peaks
from irp_b200 import _lib, ops
from irp_b200.classifier import B200Classifier, batch_stats
from irp_b200.stage import PackedImages, taps_for
from oracle import classifier_ref

ap = argparse.ArgumentParser()
ap.add_argument("--steps", type=int, default=3)
ap.add_argument("--warmup", type=int, default=3)
args = ap.parse_args()
dev = torch.device("cuda:0")
N, B, C = 27000, 256, 10
packed, ids, hw = bench.make_workload(N, seed=0, device=dev)
taps = max(taps_for(int(h), int(w), _lib.TRANSFORM_VAL_256) for h, w in hw)
packed = PackedImages(packed.pixels, packed.offsets, packed.hw, taps, packed.offsets_np, packed.hw_np)
labels = ids.to(dev, torch.int64)
model = B200Classifier(classifier_ref.build_classifier(C, seed=1234), dev, max_batch=B)
host = PackedImages(packed.pixels.cpu().pin_memory(), packed.offsets.cpu(), packed.hw.cpu(), taps, packed.offsets_np,
                    packed.hw_np)
copy_stream = torch.cuda.Stream()

def step(src, from_host):
    stats = []
    nxt = None
    def stage(lo):
        part = src.slice(lo, min(lo + B, N))
        if not from_host:
            return part
        with torch.cuda.stream(copy_stream):
            return part.to(dev)
    nxt = stage(0)
    for lo in range(0, N, B):
        if from_host:
            torch.cuda.current_stream().wait_stream(copy_stream)
        part = nxt
        if lo + B < N:
            nxt = stage(lo + B)
        x = ops.preprocess_ex(part.pixels, part.offsets, part.hw, taps, _lib.LAYOUT_NHWC4P, _lib.TRANSFORM_VAL_256)
        logits, pred = model.head(model.trunk.embed(x))
        stats.append(batch_stats(logits, labels[lo:lo + B]))
        if from_host:
            for t in (part.pixels, part.offsets, part.hw):
                t.record_stream(torch.cuda.current_stream())
    s = torch.stack(stats)
    return s.cpu() if from_host else s

def timed(src, from_host):
    for _ in range(args.warmup): step(src, from_host)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps): out = step(src, from_host)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / args.steps, out

sampler = bench.ClockSampler(0)
if sampler: sampler.start()
ms, s_dev = timed(packed, False)
if sampler: clocks = sampler.stop()
ms_e2e, s_host = timed(host, True)
assert torch.equal(s_dev.cpu(), s_host)

# CPU route of the reference on a bounded sample (same sizes/classes, host-generated pixels)
sample = 96
images, _ = bench.host_sample(sample, seed=0)
ref = classifier_ref.build_classifier(C, seed=1234)
t0 = time.perf_counter()
classifier_ref.evaluate_full(ref, classifier_ref.val_batches(images, np.arange(sample) % C, 32))
cpu_s = time.perf_counter() - t0
peaks = bench.load_peaks()
n_batches = (N + B - 1) // B
line = {
    "metric": "images/sec classify (val_transform + ResNet50 + Linear-ReLU-Linear head + cross-entropy stats)",
    "value": N / ms * 1e3, "unit": "images/s", "n_gpus": 1, "steps": args.steps, "warmup": args.warmup,
    "ms_per_step": ms, "higher_is_better": True, "dtype": "bf16", "data": "synthetic",
    "config": {"workload": "SURVEY 8f N1: 27,000 mixed-resolution uint8 images (~300x400), batch 256, 10 classes; "
                           "inputs larger than L2 (7.4 GB of pixels)"},
    "e2e": {"value": N / ms_e2e * 1e3, "unit": "images/s", "h2d_bytes_per_step": int(host.pixels.numel()),
            "d2h_bytes_per_step": int(n_batches * 3 * 8)},
    "roofline": {"bound": "tensor", "achieved": bench.FLOPS_PER_IMAGE * N / (ms * 1e-3) / 1e12,
                 "peak": peaks["tflops_sustained"], "unit": "TFLOP/s",
                 "frac": bench.FLOPS_PER_IMAGE * N / (ms * 1e-3) / 1e12 / peaks["tflops_sustained"], "traffic": None,
                 "note": "whole step (preprocess + trunk + head) against the conv FLOPs; same trunk kernels as bench.py"},
    "cpu_baseline": {"value": sample / cpu_s, "unit": "images/s", "cores": torch.get_num_threads(), "kind": "port",
                     "sample": f"{sample} images: PIL val_transform + torchvision ResNet-50 fp32 + head, batch 32"},
    "gpu_launches": n_batches * (3 + 46 + 2 + 1),
    "clocks": clocks if sampler else None,
}
print(json.dumps(line))
