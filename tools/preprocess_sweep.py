"""BASELINE.json configs[4]: preprocessing-only sweep -- center-crop/resize/normalise of mixed-resolution uint8
batches to bf16, achieved GB/s (algorithmic bytes of SURVEY.md section 8d) versus batch size."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "image-recognition-pipeline_b200")); sys.path.insert(0, ROOT)
import numpy as np, torch
import bench
from irp_b200 import _lib, ops
dev = torch.device("cuda:0")
peaks = bench.load_peaks()
packed, ids, hw = bench.make_workload(4096, seed=0, device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
rows = []
for b in (1, 8, 32, 128, 256, 512, 1024, 4096):
    part = packed.slice(0, b)
    byts = bench.algorithmic_preprocess_bytes(hw[:b])
    for layout, name in ((_lib.LAYOUT_NHWC4P, "nhwc4p"), (_lib.LAYOUT_NCHW, "nchw")):
        fn = lambda: ops.preprocess(part.pixels, part.offsets, part.hw, packed.max_taps, layout)
        for _ in range(3): fn()
        torch.cuda.synchronize(); ms = []
        for _ in range(7):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ms.append(e0.elapsed_time(e1))
        t = float(np.median(ms))
        rows.append({"batch": b, "layout": name, "ms": t, "images_per_s": b / t * 1e3, "GBps": byts / t / 1e6,
                     "frac_of_hbm": byts / t / 1e6 / peaks["hbm_gbs"]})
        print(json.dumps(rows[-1]), flush=True)
