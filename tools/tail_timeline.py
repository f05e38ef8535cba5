"""Developer tool (torchrun): per-phase timeline of the stage's multi-GPU tail (PCA + gather + LOF) on synthetic
features, 27 000 rows per rank.  Prints rank 0's CUDA-event times and the host wall time per phase."""
import os, sys, time
import numpy as np, torch, torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "image-recognition-pipeline_b200")); sys.path.insert(0, ROOT)
from irp_b200 import ops, _lib
from irp_b200.stage import OutlierStage, CudaBackend
from oracle import synth
import ctypes as C

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device(f"cuda:{local}")
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
n, d, k = 27000, 2048, 50
base = torch.from_numpy(synth.embedding_like(3000, d, seed=1 + rank)).to(dev)
feats = (base.repeat(9, 1) + 0.3 * torch.randn(n, d, device=dev)).contiguous()
ids = torch.from_numpy(synth.class_assignment(n, seed=rank)).to(dev).to(torch.int32)

class FakeTrunk:  # OutlierStage only needs .device / .max_batch from the backend's trunk here
    device, max_batch = dev, 256
backend = CudaBackend.__new__(CudaBackend); backend.trunk = FakeTrunk(); backend.device = dev
stage = OutlierStage(backend, batch_size=256, pca_components=k)

marks = []
def mark(name):
    e = torch.cuda.Event(enable_timing=True); e.record(); marks.append((name, e, time.perf_counter()))

def all_reduce(t): dist.all_reduce(t, op=dist.ReduceOp.SUM)

def lof_phases(z, group, n_groups, kk, cont, tag):
    lib = _lib.init(local)
    nn, dd = z.shape
    ws_bytes = lib.irp_lof_workspace_bytes(nn, dd, kk)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    vec = lambda: torch.empty(nn, dtype=torch.float64, device=dev)
    kdist, lrd, score = vec(), vec(), vec()
    p = lambda t: C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    mark(tag + ":alloc")
    _lib.check(lib.irp_lof_knn_part(p(z), nn, dd, p(group), n_groups, kk, rank, world, p(kdist), p(ws), ws_bytes, st), "knn"); mark(tag + ":knn_part")
    if world > 1: all_reduce(kdist); mark(tag + ":ar_kdist")
    _lib.check(lib.irp_lof_lrd_part(nn, n_groups, kk, rank, world, p(kdist), p(lrd), p(ws), ws_bytes, st), "lrd"); mark(tag + ":lrd_part")
    if world > 1: all_reduce(lrd); mark(tag + ":ar_lrd")
    _lib.check(lib.irp_lof_score_part(nn, n_groups, kk, rank, world, p(lrd), p(score), p(ws), ws_bytes, st), "score"); mark(tag + ":score_part")
    if world > 1: all_reduce(score); mark(tag + ":ar_score")
    scores, offsets, flags = vec(), torch.empty(n_groups, dtype=torch.float64, device=dev), torch.empty(nn, dtype=torch.uint8, device=dev)
    _lib.check(lib.irp_lof_finish(nn, n_groups, kk, C.c_double(cont), p(score), p(scores), p(offsets), p(flags), p(ws), ws_bytes, st), "finish"); mark(tag + ":finish")

def tail():
    marks.clear(); mark("start")
    pca = stage.fit_pca(feats); mark("fit_pca(bcast+cov+allreduce+fit)")
    z = stage.transform(feats, pca); mark("transform")
    z_all = stage.gather_rows(z).contiguous(); mark("gather z")
    ids_all = stage.gather_rows(ids).contiguous(); mark("gather ids")
    lof_phases(z_all, ids_all, 10, 30, 0.05, "class")
    lof_phases(z_all, None, 1, 75, 0.03, "global")
    torch.cuda.synchronize()

for _ in range(3):
    tail()
    if world > 1: dist.barrier()
tail()
if rank == 0:
    print(f"world {world}, rows total {n * world}")
    for (n0, e0, t0), (n1, e1, t1) in zip(marks[:-1], marks[1:]):
        print(f"  {n1:36s} gpu {e0.elapsed_time(e1):8.3f} ms   host {1e3 * (t1 - t0):8.3f} ms")
    print(f"  total gpu {marks[0][1].elapsed_time(marks[-1][1]):.3f} ms")
if world > 1: dist.destroy_process_group()
