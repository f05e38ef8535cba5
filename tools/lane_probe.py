"""Developer tool: embed-phase time (preprocess + trunk over a slice of the bench workload) with the batches
alternating between 1, 2 and 3 lanes (CudaBackend(lanes=...): one trunk handle and one stream per lane); features are
compared with the single-lane result."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "image-recognition-pipeline_b200"))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from irp_b200.stage import CudaBackend, OutlierStage, ResNet50Trunk  # noqa: E402
from oracle import stage_ref  # noqa: E402

dev = torch.device("cuda:0")
B = 256
n_img = int(os.environ.get("N_IMG", "8192"))
packed, _ids, _hw = bench.make_workload(n_img, 0, dev)
trunk = ResNet50Trunk(stage_ref.full_resnet50(1234), dev, max_batch=B)


def timed(fn, warm=1, reps=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


results, ref = [], None
for lanes in (1, 2, 3, 1, 2):
    stage = OutlierStage(CudaBackend(trunk, lanes=lanes), batch_size=B)
    f = stage.embed_packed(packed)
    torch.cuda.synchronize()
    if ref is None:
        ref = f.clone()
    diff = float((f - ref).abs().max())
    ms = timed(lambda: stage.embed_packed(packed))
    print(f"embed_packed lanes {lanes}: {ms:.2f} ms for {n_img} images ({n_img / ms:.1f} k img/s), "
          f"max|diff| vs one lane {diff:.3g}", flush=True)
    results.append({"lanes": lanes, "ms": ms, "images": n_img, "max_abs_diff": diff})
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
with open(os.path.join(ROOT, "gpurun_out", "lane_probe.jsonl"), "w") as fh:
    for r in results:
        fh.write(json.dumps(r) + "\n")
