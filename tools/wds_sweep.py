"""SURVEY.md 8f N2 measurement: the WebDataset stage's Lanczos resize_and_crop_image on mixed-resolution uint8
batches -> uint8 [n,224,224,3]; achieved GB/s on algorithmic bytes (source bytes inside the crop window +
224*224*3 output bytes) versus batch size, with the CPU route (Pillow, one thread per image like the reference
loop) timed on a sample."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "image-recognition-pipeline_b200")); sys.path.insert(0, ROOT)
import numpy as np, torch
import bench
from irp_b200 import _lib, ops
from irp_b200.stage import PackedImages, taps_for
dev = torch.device("cuda:0")
peaks = bench.load_peaks()
packed, ids, hw = bench.make_workload(4096, seed=0, device=dev)
T = _lib.TRANSFORM_WDS_LANCZOS
taps = max(taps_for(int(h), int(w), T) for h, w in hw)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

def alg_bytes(hw):
    short = np.minimum(hw[:, 0], hw[:, 1]).astype(np.float64)
    return float((3.0 * np.minimum(short, hw[:, 0]) * np.minimum(short, hw[:, 1]) + 3 * 224 * 224).sum())

for b in (1, 8, 32, 128, 256, 512, 1024, 4096):
    part = packed.slice(0, b)
    byts = alg_bytes(hw[:b])
    fn = lambda: ops.preprocess_ex(part.pixels, part.offsets, part.hw, taps, _lib.LAYOUT_U8_HWC, T)
    for _ in range(3): fn()
    torch.cuda.synchronize(); ms = []
    for _ in range(7):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ms.append(e0.elapsed_time(e1))
    t = float(np.median(ms))
    print(json.dumps({"batch": b, "layout": "u8_hwc", "filter": "lanczos", "max_taps": taps, "ms": t,
                      "images_per_s": b / t * 1e3, "GBps": byts / t / 1e6,
                      "frac_of_hbm": byts / t / 1e6 / peaks["hbm_gbs"]}), flush=True)

# CPU route on a sample: the reference function's PIL calls
from PIL import Image
images, _ = bench.host_sample(256, seed=0)
pil = [Image.fromarray(im) for im in images]
from oracle import pil_resample
t0 = time.perf_counter()
for im in pil:
    w, h = im.size
    oh, ow = pil_resample.wds_resized_size(h, w)
    r = im.resize((ow, oh), Image.Resampling.LANCZOS)
    l, tp = (ow - 224) // 2, (oh - 224) // 2
    r.crop((l, tp, l + 224, tp + 224))
cpu = time.perf_counter() - t0
print(json.dumps({"cpu_baseline": {"value": len(pil) / cpu, "unit": "images/s", "cores": 1, "kind": "port",
                                   "sample": "256 images, PIL resize(LANCZOS) + crop, single thread like the reference loop"}}))
