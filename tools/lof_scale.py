"""Developer tool: LOF cost on one rank of W when every rank holds 27k rows (weak scaling): the global pass runs the
neighbour search for 1/W of the query tiles against all W*27k rows."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "image-recognition-pipeline_b200")); sys.path.insert(0, ROOT)
import numpy as np, torch
import bench
from irp_b200 import ops
from irp_b200.stage import CudaBackend, OutlierStage, ResNet50Trunk
from oracle import stage_ref
dev = torch.device("cuda:0")
N = 27000
packed, ids, hw = bench.make_workload(N, seed=0, device=dev)
trunk = ResNet50Trunk(stage_ref.full_resnet50(seed=1234), dev, max_batch=256)
stage = OutlierStage(CudaBackend(trunk), batch_size=256, pca_components=50)
feats = stage.embed_packed(packed)
pca = stage.fit_pca(feats)
z = stage.transform(feats, pca)
idl = ids.to(dev).to(torch.int32)
def timed(fn):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); fn(); e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1)
for W in (1, 2, 4, 8):
    g = torch.Generator(device=dev).manual_seed(W)
    za = torch.cat([z + (0.05 * r) * torch.randn(z.shape, device=dev, generator=g) for r in range(W)]).contiguous()
    ia = idl.repeat(W).contiguous()
    noop = lambda t: None
    tc = timed(lambda: ops.lof_sharded(za, ia, 10, 30, 0.05, 0, W, noop))
    tg = timed(lambda: ops.lof_sharded(za, None, 1, 75, 0.03, 0, W, noop))
    print(f"W={W}: rows {za.shape[0]}: per-class LOF {tc:6.2f} ms, global LOF {tg:6.2f} ms (rank 0 of {W})", flush=True)
