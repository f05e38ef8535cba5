"""Developer tool: where one step of the stage spends its time (CUDA events + host wall clock per phase)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "image-recognition-pipeline_b200")); sys.path.insert(0, ROOT)
import torch
import bench
from irp_b200.stage import CudaBackend, OutlierStage, ResNet50Trunk, PackedImages
from oracle import stage_ref
dev = torch.device("cuda:0")
N = int(sys.argv[1]) if len(sys.argv) > 1 else 27000
packed, ids, hw = bench.make_workload(N, seed=0, device=dev)
trunk = ResNet50Trunk(stage_ref.full_resnet50(seed=1234), dev, max_batch=256)
stage = OutlierStage(CudaBackend(trunk), batch_size=256, pca_components=50)
pin = torch.empty(packed.pixels.numel(), dtype=torch.uint8, pin_memory=True); pin.copy_(packed.pixels)
host = PackedImages(pin, packed.offsets.cpu().pin_memory(), packed.hw.cpu().pin_memory(), packed.max_taps, packed.offsets_np, packed.hw_np)

def timed(from_host):
    src = host if from_host else packed
    marks = []
    def mark(name):
        e = torch.cuda.Event(enable_timing=True); e.record(); marks.append((name, e, time.perf_counter()))
    torch.cuda.synchronize(); mark("start")
    feats = stage.embed_packed(src, from_host=from_host); mark("embed")
    pca = stage.fit_pca(feats); mark("fit_pca")
    z = stage.transform(feats, pca); mark("transform")
    idl = ids.to(dev, non_blocking=True).to(torch.int32)
    za, ia = stage.gather_rows(z), stage.gather_rows(idl); mark("gather")
    out = stage.detect(za.contiguous(), ia.contiguous(), 10); mark("detect")
    if from_host:
        h = (feats.cpu(), za.cpu(), out[0].cpu(), out[1].cpu()); mark("d2h")
    torch.cuda.synchronize(); t_end = time.perf_counter()
    line = []
    for (n0, e0, w0), (n1, e1, w1) in zip(marks[:-1], marks[1:]):
        line.append(f"{n1} gpu {e0.elapsed_time(e1):7.1f} ms host {1e3*(w1-w0):7.1f} ms")
    print(("host " if from_host else "hbm  ") + " | ".join(line) + f" | total wall {1e3*(t_end-marks[0][2]):.1f} ms")
for _ in range(2): timed(False)
timed(False); timed(False)
timed(True); timed(True); timed(True)
