"""Benchmark of the B200 outlier-detection stage (BASELINE.json: images/sec for embed + PCA + outlier-score).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

One "step" = one pass of the whole hot path over the workload of BASELINE.json configs[1]: 27 000 synthetic
mixed-resolution uint8 images (~300x400) -> fused preprocess -> ResNet-50 embeddings (batch 256, bf16) -> PCA(50)
-> per-class + global LOF.  With N GPUs every rank processes its own 27 000 images (weak scaling); the PCA
partial sums are all-reduced once and the projected rows gathered for scoring.

Rank 0 prints ONE JSON line.  `value` is measured with the inputs resident in HBM; `e2e` runs the same step from
pinned host buffers (H2D of every batch and D2H of features / projection / flags inside the timed region).
`--impl reference` times the reference's own torch-CPU route (oracle/stage_ref.py, the same torchvision / sklearn
calls) on the host cores for a bounded sample per step.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "image-recognition-pipeline_b200")
for _p in (ROOT, PKG):
    if _p not in sys.path:
        sys.path.insert(0, _p)


def host_cores() -> int:
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:  # pragma: no cover
        return os.cpu_count() or 1


def _is_reference_arm(argv) -> bool:
    for i, a in enumerate(argv):
        if a == "--impl" and i + 1 < len(argv) and argv[i + 1] == "reference":
            return True
        if a == "--impl=reference":
            return True
    return False


if _is_reference_arm(sys.argv):
    # torch.distributed.run exports OMP_NUM_THREADS=1 to every rank; the CPU arm must use all the host cores whatever
    # launched it, so the thread pools are sized BEFORE torch / numpy / sklearn are imported.
    for _v in ("OMP_NUM_THREADS", "MKL_NUM_THREADS", "OPENBLAS_NUM_THREADS", "NUMEXPR_NUM_THREADS"):
        os.environ[_v] = str(host_cores())

import numpy as np  # noqa: E402
import torch  # noqa: E402

METRIC = "images/sec embed+PCA+outlier-score"
UNIT = "images/s"
N_IMAGES = 27000
BATCH = 256
PCA_K = 50
N_CLASSES = 10
FLOPS_PER_IMAGE = 8.1743e9  # 53 convolutions of the ResNet-50 trunk
WORKLOAD = ("configs[1]: Animals-10-sized synthetic set, 27,000 mixed-resolution uint8 images (~300x400) per GPU -> "
            "preprocess, ResNet50 embed (bf16, batch 256), PCA(50), LOF per-class(k=30,5%) + global(k=75,3%)")


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return {"hbm_gbs": p["hbm_gbs"], "tflops_sustained": p.get("bf16_tflops_sustained", p["bf16_tflops"]),
                "tflops_burst": p["bf16_tflops"], "source": "measured"}
    return {"hbm_gbs": 6650.0, "tflops_sustained": 1400.0, "tflops_burst": 1590.0, "source": "fallback"}


# --------------------------------------------------------------------------------------------------------------
# clocks sampling (nvidia-smi during the timed region)
# --------------------------------------------------------------------------------------------------------------
class ClockSampler:
    """SM clock and throttle reasons of one GPU, sampled every 200 ms DURING the timed region.

    Reads NVML in-process (pynvml: the library behind nvidia-smi).  An `nvidia-smi -lms` child process, which this
    used to be, re-initialises NVML over every GPU of the box and stalled one timed step of a 4-GPU run by
    100-300 ms about two seconds after its start (reproduced three times; gone with the sampler off); it remains
    only as the fallback when pynvml is missing.  Only samples taken after `begin()` are reported."""
    QUERY = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

    def __init__(self, index):
        self.index = index
        self.rows = []          # (time, sm_mhz, max_mhz, set of reasons)
        self.proc = None
        self.thread = None
        self.stop_flag = threading.Event()
        self.t_begin = 0.0
        self.source = None

    def begin(self):
        self.t_begin = time.time()

    # ---- NVML in-process ----
    def _nvml_handle(self):
        import pynvml
        pynvml.nvmlInit()
        try:
            p = torch.cuda.get_device_properties(self.index)
            bus = "%08X:%02X:%02X.0" % (p.pci_domain_id, p.pci_bus_id, p.pci_device_id)
            return pynvml, pynvml.nvmlDeviceGetHandleByPciBusId(bus.encode())
        except Exception:
            vis = os.environ.get("CUDA_VISIBLE_DEVICES", "")
            ids = [v for v in vis.split(",") if v.strip().isdigit()]
            phys = int(ids[self.index]) if self.index < len(ids) else self.index
            return pynvml, pynvml.nvmlDeviceGetHandleByIndex(phys)

    def _poll_nvml(self, pynvml, handle):
        bits = {"sw_power_cap": 0x4, "hw_slowdown": 0x8, "sw_thermal_slowdown": 0x20, "hw_thermal_slowdown": 0x40}
        mx = float(pynvml.nvmlDeviceGetMaxClockInfo(handle, pynvml.NVML_CLOCK_SM))
        while not self.stop_flag.is_set():
            try:
                sm = float(pynvml.nvmlDeviceGetClockInfo(handle, pynvml.NVML_CLOCK_SM))
                mask = int(pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(handle))
                self.rows.append((time.time(), sm, mx, {n for n, b in bits.items() if mask & b}))
            except Exception:
                pass
            self.stop_flag.wait(0.2)

    # ---- nvidia-smi fallback ----
    def _read_smi(self):
        for line in self.proc.stdout:
            r = [p.strip() for p in line.split(",")]
            if len(r) >= 7:
                try:
                    reasons = {n for n, v in zip(self.NAMES, r[3:7]) if v.lower().startswith("active")}
                    self.rows.append((time.time(), float(r[0]), float(r[1]), reasons))
                except ValueError:
                    continue

    def start(self):
        try:
            pynvml, handle = self._nvml_handle()
            self.source = "nvml"
            self.thread = threading.Thread(target=self._poll_nvml, args=(pynvml, handle), daemon=True)
            self.thread.start()
            return
        except Exception:
            pass
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.QUERY}",
                                          "--format=csv,noheader,nounits", "-lms", "200"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
        except OSError:
            return
        self.source = "nvidia-smi"
        self.thread = threading.Thread(target=self._read_smi, daemon=True)
        self.thread.start()

    def stop(self):
        if self.source is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvml / nvidia-smi unavailable"]}
        self.stop_flag.set()
        if self.proc is not None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except subprocess.TimeoutExpired:
                self.proc.kill()
        if self.thread is not None:
            self.thread.join(timeout=2)
        rows = [r for r in self.rows if r[0] >= self.t_begin]
        sm = [r[1] for r in rows]
        mx = [r[2] for r in rows]
        reasons = set().union(*[r[3] for r in rows]) if rows else set()
        busy = [s for s in sm if s > 0]
        return {"sm_mhz": statistics.median(busy) if busy else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "source": self.source}


# --------------------------------------------------------------------------------------------------------------
# synthetic workload
# --------------------------------------------------------------------------------------------------------------
def make_workload(n, seed, device):
    """Packed uint8 HWC images generated ON THE DEVICE (smooth 8x8 base upsampled + noise + class tint), class ids.
    Sizes follow oracle/synth.py (mixed resolution around 300x400, ~45 % with a side < 224)."""
    from irp_b200.stage import ALIGN, PackedImages, taps_for
    from oracle import synth

    hw = synth.mixed_resolution_sizes(n, seed=seed)
    ids = synth.class_assignment(n, seed=seed)
    nbytes = hw[:, 0].astype(np.int64) * hw[:, 1] * 3
    padded = (nbytes + ALIGN - 1) // ALIGN * ALIGN
    offsets = np.concatenate([[0], np.cumsum(padded)[:-1]]).astype(np.int64)
    total = int(padded.sum())
    pixels = torch.empty(total, dtype=torch.uint8, device=device)
    g = torch.Generator(device=device).manual_seed(seed)
    for i in range(n):
        h, w = int(hw[i, 0]), int(hw[i, 1])
        base = torch.rand(1, 3, 8, 8, device=device, generator=g) * 255.0
        img = torch.nn.functional.interpolate(base, size=(h, w), mode="bilinear", align_corners=False)[0]
        img = img + torch.randn(3, h, w, device=device, generator=g) * 12.0 + 5.0 * float(ids[i])
        pixels[offsets[i]:offsets[i] + nbytes[i]] = img.clamp_(0, 255).permute(1, 2, 0).reshape(-1).to(torch.uint8)
    taps = max(taps_for(int(h), int(w)) for h, w in hw)
    packed = PackedImages(pixels, torch.from_numpy(offsets).to(device), torch.from_numpy(hw).to(device), taps,
                          offsets, hw)
    return packed, torch.from_numpy(ids), hw


def algorithmic_preprocess_bytes(hw):
    """source bytes inside the crop window + 3*224*224*2 output bytes per image."""
    short = np.minimum(hw[:, 0], hw[:, 1]).astype(np.float64)
    window = 3.0 * np.ceil(224.0 / 232.0 * short) ** 2
    return float((window + 3 * 224 * 224 * 2).sum())


# --------------------------------------------------------------------------------------------------------------
# reference arm (CPU)
# --------------------------------------------------------------------------------------------------------------
def cpu_stage_sample(images, labels, k):
    """The reference route on the host for a bounded sample: embed (batch 32) + PCA + detect_outliers."""
    import warnings

    from oracle import stage_ref

    t0 = time.perf_counter()
    feats = stage_ref.embed_arrays(images, batch_size=32, seed=1234)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        z, _ = stage_ref.pca_default(feats, min(k, len(images) - 1))
        stage_ref.detect_outliers(z, labels)
    return time.perf_counter() - t0


def host_sample(n, seed):
    from oracle import synth

    rng = np.random.default_rng(seed)
    hw = synth.mixed_resolution_sizes(n, seed=seed)
    ids = synth.class_assignment(n, seed=seed)
    images = [synth.smooth_image(rng, int(h), int(w), int(c)) for (h, w), c in zip(hw, ids)]
    return images, np.array([f"class{c}" for c in ids])


def run_reference(args, rank, world):
    if rank != 0:
        return
    sample = 192
    images, labels = host_sample(sample, seed=0)
    cores = torch.get_num_threads()
    warm_labels = np.array([f"class{i % 2}" for i in range(32)])  # two classes of 16: LOF needs >= 2 rows each
    for _ in range(args.warmup):
        cpu_stage_sample(images[:32], warm_labels, PCA_K)
    times = [cpu_stage_sample(images, labels, PCA_K) for _ in range(args.steps)]
    sec = sum(times) / len(times)
    value = sample / sec
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "sample_per_step": sample, "batch": 32, "pca_components": PCA_K},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{sample} images of the same workload per step: PIL transform + torchvision "
                                   f"ResNet-50 fp32 (batch 32) + sklearn PCA + LocalOutlierFactor on {cores} threads "
                                   "(oracle/stage_ref.py restates functions/data_curation.py:654-728 with the same "
                                   "library calls; /root/reference is not present on the GPU box)"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------------------------------
# our arm
# --------------------------------------------------------------------------------------------------------------
class TimedBackend:
    """CudaBackend that brackets the preprocess and trunk launches with CUDA events on the launching stream."""

    def __init__(self, inner):
        self.inner = inner
        self.device = inner.device
        self.trunk = inner.trunk
        self.events = []
        self.enabled = False

    def __getattr__(self, name):
        # everything that is not timed here (cov_accumulate, pca_fit, pca_transform, lof, lof_sharded, ...) is the
        # product backend's own method -- in particular lof_sharded, without which a multi-rank run would fall back
        # to every rank searching all rows
        return getattr(self.inner, name)

    def embed(self, part, max_taps):
        from irp_b200 import _lib, ops
        if not self.enabled:
            return self.inner.embed(part, max_taps)
        e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        e[0].record()
        x = ops.preprocess(part.pixels, part.offsets, part.hw, max_taps, _lib.LAYOUT_NHWC4P)
        e[1].record()
        out = self.trunk.embed(x)
        e[2].record()
        self.events.append((e, len(part)))
        return out


# DRAM bytes (dram__bytes_read.sum + dram__bytes_write.sum) of the 46 kernels of ONE irp_resnet50_embed call at batch 256,
# from the ncu pass committed as profiles/r01_ncu_trunk_traffic_v6.csv (5 411 MB read + 3 282 MB written; the pass
# captured 49 launches across two calls, profiles/r01_trunk_layer_table_v6.md lists the 46 of one call).
TRUNK_DRAM_BYTES_PER_CALL = 8.694e9


def launches_per_step(n_images, batch, dim, max_taps):
    """Kernels of libirp_b200.so launched per step (counted from the launch sites in csrc/*.cu)."""
    batches = (n_images + batch - 1) // batch
    pre = 3 + (1 if max_taps > 6 else 0)  # resample_plan + horizontal pass + vertical pass (+ many-tap path)
    trunk = 46                   # stem+pool, 16 3x3, 3 downsample, 9 conv1, 9 conv3, 7 chained conv3+conv1 (the first
                                 # with the layer1 shortcut conv folded in), avgpool
    per_batch = pre + trunk
    cov = 3                      # split_transpose, add_count, cov_gemm
    k = PCA_K
    lanczos_steps = max(3 * k + 10, 96)  # first convergence check; the bench data converges there
    # assemble, start vector, 5 kernels per step, close, bisect, inverse iteration, mgs, residual, Ritz, sign, clip
    fit = 2 + 5 * lanczos_steps + 1 + 4 + 2 + 1 if dim >= 512 and 8 * k <= dim else 1 + (dim - 1) + 6
    transform = 1
    # iota/count/scan/scatter (grouped only), sort key + radix sort (4 CUB passes + histogram), sqnorm, gather,
    # knn, lrd, score, unsort, percentile, flag
    lof_global = 1 + 1 + 5 + 8
    lof_grouped = 3 + lof_global
    return batches * per_batch + cov + fit + transform + lof_grouped + lof_global


def run_ours(args, rank, local_rank, world):
    import torch.distributed as dist

    from irp_b200.stage import CudaBackend, OutlierStage, ResNet50Trunk
    from oracle import stage_ref  # weights only: the seeded random-init torchvision module (no network)

    assert torch.cuda.is_available(), "bench.py needs a GPU (the CUDA path has no CPU fallback)"
    torch.cuda.set_device(local_rank)
    dev = torch.device(f"cuda:{local_rank}")
    if world > 1 and not dist.is_initialized():
        dist.init_process_group("nccl", device_id=dev)
    peaks = load_peaks()

    packed, ids, hw = make_workload(N_IMAGES, seed=rank, device=dev)
    trunk = ResNet50Trunk(stage_ref.full_resnet50(seed=1234), dev, max_batch=BATCH)
    backend = TimedBackend(CudaBackend(trunk))
    stage = OutlierStage(backend, batch_size=BATCH, pca_components=PCA_K)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step(from_host=False, src=None):
        res = stage.run(src if src is not None else packed, ids, N_CLASSES, from_host=from_host)
        if from_host:  # what process_image_directory / detect_outliers hand back to the host
            out = (stage.to_host(res.features, "features"), stage.to_host(res.z[: len(packed)], "z"),
                   stage.to_host(res.class_outliers, "class_flags"), stage.to_host(res.global_outliers, "global_flags"))
            return res, out
        # Every step ends with the host waiting for its result (the drop-in returns the flags to the caller).  Without
        # it the host runs a whole step ahead of the GPU; with 4 ranks one of the first timed steps then stalled for
        # 100-300 ms (ranks enqueue their collectives out of phase) -- with the step closed like this it does not.
        torch.cuda.current_stream().synchronize()
        return res, None

    # ---- value: inputs resident in HBM ----
    sampler = ClockSampler(local_rank)
    if rank == 0 and os.environ.get("IRP_BENCH_NO_SAMPLER", "0") == "0":
        sampler.start()
    for _ in range(args.warmup):
        step()
    barrier()
    backend.enabled = os.environ.get("IRP_BENCH_NO_EVENTS", "0") == "0"
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    sampler.begin()
    e0.record()
    step_marks = [e0]
    for _ in range(args.steps):
        res, _ = step()
        m = torch.cuda.Event(enable_timing=True)
        m.record()
        step_marks.append(m)
    e1.record()
    barrier()
    backend.enabled = False
    per_step_ms = [round(a.elapsed_time(b), 2) for a, b in zip(step_marks[:-1], step_marks[1:])]
    if getattr(stage, "traces", None):
        for i, tr in enumerate(stage.traces):
            print(f"[trace rank {rank}] step {i}: {tr}", file=sys.stderr, flush=True)
    clocks = sampler.stop() if rank == 0 else None
    ms_total = e0.elapsed_time(e1)
    t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_step = t.item() / args.steps
    n_total = N_IMAGES * world
    value = n_total / (ms_step * 1e-3)

    pre_ms = sum(e[0].elapsed_time(e[1]) for e, _ in backend.events)
    emb_ms = sum(e[1].elapsed_time(e[2]) for e, _ in backend.events)
    # per-rank embed time (preprocess + trunk): the step ends when the SLOWEST rank reaches the PCA all-reduce, so
    # the spread between GPUs of one box shows up as waiting time in every other rank's "pca_lof_other"
    rank_embed_ms = None
    if world > 1:
        mine = torch.tensor([(pre_ms + emb_ms) / args.steps], dtype=torch.float64, device=dev)
        allr = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(allr, mine)
        rank_embed_ms = [round(float(t.item()), 3) for t in allr]
    n_emb = sum(n for _, n in backend.events)
    calls = len(backend.events)
    conv_tflops = FLOPS_PER_IMAGE * n_emb / (max(emb_ms, 1e-9) * 1e-3) / 1e12
    pre_gbs = algorithmic_preprocess_bytes(hw) * args.steps / (max(pre_ms, 1e-9) * 1e-3) / 1e9

    # ---- e2e: pinned host buffers, copies inside the timed region ----
    host = None
    e2e = None
    try:
        from irp_b200.stage import PackedImages
        pin = torch.empty(packed.pixels.numel(), dtype=torch.uint8, pin_memory=True)
        pin.copy_(packed.pixels)
        host = PackedImages(pin, packed.offsets.cpu().pin_memory(), packed.hw.cpu().pin_memory(), packed.max_taps,
                            packed.offsets_np, packed.hw_np)
    except RuntimeError as ex:  # not enough pinnable memory
        e2e = {"value": None, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0, "error": str(ex)[:120]}
    if host is not None:
        step(True, host)
        barrier()
        t0 = time.perf_counter()
        e0.record()
        for _ in range(args.steps):
            _, out = step(True, host)
        e1.record()
        barrier()
        wall = time.perf_counter() - t0
        t = torch.tensor([max(e0.elapsed_time(e1) * 1e-3, wall)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        h2d = int(pin.numel() + host.offsets.numel() * 8 + host.hw.numel() * 4 + ids.numel() * 4)
        d2h = int(sum(o.numel() * o.element_size() for o in out))
        e2e = {"value": n_total / (t.item() / args.steps), "unit": UNIT, "h2d_bytes_per_step": h2d,
               "d2h_bytes_per_step": d2h}

    if rank != 0:
        return
    # ---- CPU baseline on a bounded sample (rank 0, N=1 only) ----
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        sample = 384
        images, labels = host_sample(sample, seed=0)
        sec = cpu_stage_sample(images, labels, PCA_K)
        cpu = {"value": sample / sec, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
               "sample": f"{sample} images of the same workload, once: PIL transform + torchvision ResNet-50 fp32 "
                         f"(batch 32) + sklearn PCA + LocalOutlierFactor (oracle/stage_ref.py) in {sec:.1f} s"}
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16",
        "data": "synthetic",
        "config": {"workload": WORKLOAD, "images_per_gpu": N_IMAGES, "batch": BATCH, "pca_components": PCA_K,
                   "classes": N_CLASSES, "weights": "random-init torchvision resnet50 (seed 1234)",
                   "l2": f"inputs ({packed.pixels.numel() / 1e9:.1f} GB per GPU) and activations larger than L2; "
                         "no explicit flush",
                   "parallelism": f"dp{world}: images sharded, one all-reduce of the PCA partial sums"},
        "roofline": {"bound": "tensor", "achieved": conv_tflops, "peak": peaks["tflops_sustained"], "unit": "TFLOP/s",
                     "frac": conv_tflops / peaks["tflops_sustained"], "traffic": TRUNK_DRAM_BYTES_PER_CALL,
                     "traffic_source": "ncu dram__bytes_read+write over the 46 kernels of one trunk call, "
                                       "profiles/r01_ncu_trunk_traffic_v6.csv",
                     "kernel": "conv_gemm2_kernel / conv_chain_kernel / conv3x3_c64_kernel / stem_pool_kernel (the 53 "
                               "convolutions of one irp_resnet50_embed call, batch 256)",
                     "per_launch": f"{FLOPS_PER_IMAGE:.4e} FLOP/image x {BATCH} images per trunk call; "
                                   f"{calls} calls timed with CUDA events, mean {emb_ms / max(calls, 1):.3f} ms",
                     "peak_source": f"{peaks['source']} bf16_tflops_sustained"},
        "roofline_preprocess": {"bound": "hbm", "achieved": pre_gbs, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                                "frac": pre_gbs / peaks["hbm_gbs"], "traffic": None,
                                "kernel": "hpass_kernel + vpass_kernel (+ resample_plan_kernel)",
                                "per_launch": "3*ceil(224/232*short)^2 source bytes + 301056 output bytes per image"},
        "stage_ms": {"preprocess": pre_ms / args.steps, "trunk": emb_ms / args.steps,
                     "pca_lof_other": ms_step - (pre_ms + emb_ms) / args.steps,
                     "embed_per_rank": rank_embed_ms, "per_step": per_step_ms},
        "cpu_baseline": cpu, "e2e": e2e, "clocks": clocks,
        "gpu_launches": launches_per_step(N_IMAGES, BATCH, 2048, packed.max_taps) * args.steps,
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if world != args.gpus and world == 1 and args.gpus > 1:
        raise SystemExit("launch multi-GPU runs with torch.distributed.run (one rank per GPU)")
    run_ours(args, rank, local_rank, world)


if __name__ == "__main__":
    main()
