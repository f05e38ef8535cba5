"""Benchmark of the B200 outlier-detection stage (BASELINE.json: images/sec for embed + PCA + outlier-score).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config cfg2|cfg3_strong|cfg4]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

One "step" = one pass of the whole hot path over the workload.  The default (--config cfg2) is BASELINE.json
configs[1]: 27 000 synthetic mixed-resolution uint8 images (~300x400, with the rare >= 1500 px ones) -> fused
preprocess -> ResNet-50 embeddings (batch 256, bf16) -> PCA(50) -> per-class + global LOF.  With N GPUs every rank
processes its own 27 000 images (weak scaling); the PCA partial sums are all-reduced once and the projected rows
gathered for scoring.  --config cfg3_strong is configs[2] (the SAME 27 000 images sharded over the N ranks, strong
scaling); --config cfg4 is configs[3] (125 000 synthetic 224x224 images per GPU = 1 M on 8 GPUs, batch 512,
PCA(128), global scorer only).  Every line also carries the per-phase times of one extra traced step.

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
N_CLASSES = 10
FLOPS_PER_IMAGE = 8.1743e9  # 53 convolutions of the ResNet-50 trunk


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
    """SM clock and throttle reasons of one GPU, sampled every 250 ms DURING the timed region.

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
            self.stop_flag.wait(0.25)

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
# workload configurations
# --------------------------------------------------------------------------------------------------------------
CONFIGS = {
    # BASELINE.json configs[1] (the line the driver reads): weak scaling, 27 000 images per GPU
    "cfg2": dict(images_per_gpu=27000, total=None, batch=256, k=50, class_scoring=True, sizes="mixed",
                 scaling="weak",
                 workload="configs[1]: Animals-10-sized synthetic set, 27,000 mixed-resolution uint8 images (~300x400, "
                          "42 % with a side < 224, 0.2 % of 1500-2600 px) per GPU -> preprocess, ResNet50 embed (bf16, "
                          "batch 256), PCA(50), LOF per-class(k=30,5%) + global(k=75,3%)"),
    # configs[2]: the same 27 000 images sharded over the ranks
    "cfg3_strong": dict(images_per_gpu=None, total=27000, batch=256, k=50, class_scoring=True, sizes="mixed",
                        scaling="strong",
                        workload="configs[2]: the 27,000-image Animals-10-sized set of configs[1] sharded over the GPUs "
                                 "(contiguous ranges), one all-reduce of the PCA partial sums, PCA(50), LOF per-class + "
                                 "global"),
    # configs[3]: 1 M images on 8 GPUs = 125 000 per GPU
    "cfg4": dict(images_per_gpu=125000, total=None, batch=512, k=128, class_scoring=False, sizes="224",
                 scaling="weak",
                 workload="configs[3]: scale-out, 125,000 synthetic 224x224 uint8 images per GPU (1 M on 8 GPUs) -> "
                          "preprocess, ResNet50 embed (bf16, batch 512), PCA(128), global LOF(k=75,3%) only"),
}


def workload_sizes(cfg, n, seed):
    from oracle import synth
    if cfg["sizes"] == "224":
        return np.full((n, 2), 224, np.int32), synth.class_assignment(n, seed=seed)
    return synth.mixed_resolution_sizes(n, seed=seed), synth.class_assignment(n, seed=seed)


def make_workload(n, seed, device, cfg=None, lo=0, hi=None):
    """Packed uint8 HWC images generated ON THE DEVICE (smooth 8x8 base upsampled + noise + class tint), class ids.
    Sizes follow oracle/synth.py (mixed resolution around 300x400, ~42 % with a side < 224, 0.2 % of 1500-2600 px)
    or are all 224x224 (configs[3]).  [lo, hi) selects this rank's contiguous share of the n images."""
    from irp_b200.stage import ALIGN, PackedImages, taps_for

    cfg = cfg or CONFIGS["cfg2"]
    hw, ids = workload_sizes(cfg, n, seed)
    hi = n if hi is None else hi
    hw, ids = hw[lo:hi], ids[lo:hi]
    m = hi - lo
    nbytes = hw[:, 0].astype(np.int64) * hw[:, 1] * 3
    padded = (nbytes + ALIGN - 1) // ALIGN * ALIGN
    offsets = np.concatenate([[0], np.cumsum(padded)[:-1]]).astype(np.int64)
    total = int(padded.sum())
    pixels = torch.empty(total, dtype=torch.uint8, device=device)
    g = torch.Generator(device=device).manual_seed(seed * 1000003 + lo)
    if cfg["sizes"] == "224":
        ids_t = torch.from_numpy(ids).to(device)
        per = 224 * 224 * 3
        view = pixels.view(m, per)
        for s in range(0, m, 1024):
            e = min(m, s + 1024)
            base = torch.rand(e - s, 3, 8, 8, device=device, generator=g) * 255.0
            img = torch.nn.functional.interpolate(base, size=(224, 224), mode="bilinear", align_corners=False)
            img = img + torch.randn(e - s, 3, 224, 224, device=device, generator=g) * 12.0
            img = img + 5.0 * ids_t[s:e].float().view(-1, 1, 1, 1)
            view[s:e] = img.clamp_(0, 255).permute(0, 2, 3, 1).reshape(e - s, per).to(torch.uint8)
    else:
        for i in range(m):
            h, w = int(hw[i, 0]), int(hw[i, 1])
            base = torch.rand(1, 3, 8, 8, device=device, generator=g) * 255.0
            img = torch.nn.functional.interpolate(base, size=(h, w), mode="bilinear", align_corners=False)[0]
            img = img + torch.randn(3, h, w, device=device, generator=g) * 12.0 + 5.0 * float(ids[i])
            pixels[offsets[i]:offsets[i] + nbytes[i]] = img.clamp_(0, 255).permute(1, 2, 0).reshape(-1).to(torch.uint8)
    taps = max(taps_for(int(h), int(w)) for h, w in np.unique(hw, axis=0))
    packed = PackedImages(pixels, torch.from_numpy(offsets).to(device), torch.from_numpy(hw).to(device), taps,
                          offsets, hw)
    return packed, torch.from_numpy(ids), hw


def algorithmic_preprocess_bytes(hw):
    """source bytes inside the crop window + 3*224*224*2 output bytes per image."""
    short = np.minimum(hw[:, 0], hw[:, 1]).astype(np.float64)
    window = 3.0 * np.ceil(224.0 / 232.0 * short) ** 2
    return float((window + 3 * 224 * 224 * 2).sum())


def config_dict(cfg_name, world):
    """The `config` object of the JSON line; the reference arm prints the SAME object for the same --config / N."""
    cfg = CONFIGS[cfg_name]
    n_total = cfg["total"] if cfg["total"] is not None else cfg["images_per_gpu"] * world
    return {"workload": cfg["workload"], "name": cfg_name, "images_total": n_total,
            "images_per_gpu": cfg["images_per_gpu"] if cfg["images_per_gpu"] is not None else f"{n_total}/N (contiguous)",
            "batch": cfg["batch"], "pca_components": cfg["k"], "classes": N_CLASSES,
            "scorers": "per-class LOF + global LOF" if cfg["class_scoring"] else "global LOF",
            "weights": "random-init torchvision resnet50 (seed 1234)",
            "l2": "inputs (GBs per GPU) and activations larger than the 126 MB L2; no explicit flush",
            "parallelism": f"dp{world}: images sharded, one all-reduce of the PCA partial sums, all-gather of the "
                           "projected rows, three all-reduces of the sharded LOF"}


# --------------------------------------------------------------------------------------------------------------
# reference arm (CPU)
# --------------------------------------------------------------------------------------------------------------
def cpu_stage_sample(images, labels, k, class_scoring=True):
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


def host_sample(n, seed, cfg=None):
    from oracle import synth

    cfg = cfg or CONFIGS["cfg2"]
    rng = np.random.default_rng(seed)
    hw, ids = workload_sizes(cfg, n, seed)
    hw = np.minimum(hw, 1400)  # host sample only: a 2600 px synthetic image costs seconds to GENERATE on the CPU
    images = [synth.smooth_image(rng, int(h), int(w), int(c)) for (h, w), c in zip(hw, ids)]
    return images, np.array([f"class{c}" for c in ids])


REFERENCE_SAMPLE = 1024  # images per step of the CPU arm (time a >= 1 024-image sample)


def run_reference(args, rank, world):
    """The reference's CPU route (oracle/stage_ref.py: the same torchvision / sklearn calls as
    functions/data_curation.py:654-728) on ALL host cores, whatever launched this process; rank 0 only."""
    if rank != 0:
        return
    cores = host_cores()
    torch.set_num_threads(cores)
    cfg = CONFIGS[args.config]
    sample = REFERENCE_SAMPLE
    images, labels = host_sample(sample, seed=0, cfg=cfg)
    warm_labels = np.array([f"class{i % 2}" for i in range(32)])  # two classes of 16: LOF needs >= 2 rows each
    for _ in range(min(args.warmup, 2)):
        cpu_stage_sample(images[:32], warm_labels, cfg["k"])
    times = [cpu_stage_sample(images, labels, cfg["k"]) for _ in range(args.steps)]
    sec = sum(times) / len(times)
    value = sample / sec
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True,
        "scaling": cfg["scaling"], "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": config_dict(args.config, world),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
                         "host_cores": cores,
                         "sample": f"{sample} images of the same workload per step (warm-up steps: 32 images): PIL "
                                   f"transform + torchvision ResNet-50 fp32 (batch 32) + sklearn PCA({cfg['k']}) + "
                                   f"LocalOutlierFactor on {sample} rows, {torch.get_num_threads()} threads "
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
        # everything that is not timed here (cov_accumulate, pca_fit, pca_transform, lof, lof_sharded_multi, ...) is
        # the product backend's own method -- in particular the sharded LOF, without which a multi-rank run would fall
        # back to every rank searching all rows
        return getattr(self.inner, name)

    def embed(self, part, max_taps, lane=None):
        """Events go to the CURRENT stream, i.e. the lane's stream when the stage runs its batches on two lanes."""
        from irp_b200 import _lib, ops
        if not self.enabled:
            return self.inner.embed(part, max_taps, lane=lane)
        e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        e[0].record()
        x = ops.preprocess(part.pixels, part.offsets, part.hw, max_taps, _lib.LAYOUT_NHWC4P)
        e[1].record()
        out = (self.trunk if lane is None else self.inner.lane_trunks[lane]).embed(x)
        e[2].record()
        self.events.append((e, len(part)))
        return out


def union_ms(intervals):
    """Total length of the union of [a, b] intervals (ms)."""
    total, end = 0.0, None
    for a, b in sorted(intervals):
        if end is None or a > end:
            total += b - a
            end = b
        elif b > end:
            total += b - end
            end = b
    return total


def bind_near_gpu(index):
    """Pin this process to the CPUs NVML reports as closest to GPU `index` (one process per GPU): the pinned host
    buffers of the e2e leg are then first-touched, hence allocated, on the GPU's own NUMA node.  With eight ranks on
    one box the H2D copies (77 MB per 3 ms batch and GPU) otherwise cross the socket interconnect.  Best effort."""
    try:
        import pynvml
        pynvml.nvmlInit()
        p = torch.cuda.get_device_properties(index)
        bus = "%08X:%02X:%02X.0" % (p.pci_domain_id, p.pci_bus_id, p.pci_device_id)
        h = pynvml.nvmlDeviceGetHandleByPciBusId(bus.encode())
        n_words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, n_words)
        cpus = {64 * w + b for w, word in enumerate(mask) for b in range(64) if (int(word) >> b) & 1}
        allowed = cpus & set(os.sched_getaffinity(0))
        if allowed:
            os.sched_setaffinity(0, allowed)
            return len(allowed)
    except Exception:
        pass
    return None


def measured_traffic(name):
    """DRAM bytes per launch group from the committed ncu summary (profiles/r02_traffic.json, written by
    tools/traffic_summary.py from an ncu dram__bytes_read.sum + dram__bytes_write.sum pass); None if absent."""
    path = os.path.join(ROOT, "profiles", "r02_traffic.json")
    if not os.path.exists(path):
        return None, None
    with open(path) as f:
        t = json.load(f)
    e = t.get(name)
    return (e["bytes"], e["source"]) if e else (None, None)


def launches_per_step(n_images, batch, dim, k, class_scoring):
    """Kernels of libirp_b200.so launched per step (counted from the launch sites in csrc/*.cu)."""
    batches = (n_images + batch - 1) // batch
    pre = 2                      # resample_plan (taps + band schedule) + resample_fused
    trunk = 45                   # stem+pool, 16 3x3, 3 downsample, 9 conv1, 8 conv3, 7 chained conv3+conv1 (the first
                                 # with the layer1 shortcut conv folded in), last conv3 + residual + average pool
    per_batch = pre + trunk
    cov = 4                      # split_transpose, col_sum_reduce, add_count, cov_gemm
    lanczos_steps = max(3 * k + 10, 96)  # first convergence check; the bench data converges there
    # assemble, trace, start vector, 5 kernels per step, close, bisect, inverse iteration, mgs, residual, Ritz, sign, clip
    fit = 3 + 5 * lanczos_steps + 1 + 4 + 2 + 1 if dim >= 512 and 8 * k <= dim else 2 + (dim - 1) + 6
    transform = 2                # projection operands + tcgen05 projection GEMM
    # iota/count/scan/scatter (grouped only), sort key + radix sort (4 CUB passes + histogram), sqnorm, gather,
    # knn, lrd, score, unsort, percentile, flag
    lof_global = 1 + 1 + 5 + 8
    lof_grouped = 3 + lof_global
    return batches * per_batch + cov + fit + transform + (lof_grouped if class_scoring else 0) + lof_global


def run_ours(args, rank, local_rank, world):
    import torch.distributed as dist

    from irp_b200.stage import CudaBackend, OutlierStage, ResNet50Trunk, shard_range
    from oracle import stage_ref  # weights only: the seeded random-init torchvision module (no network)

    assert torch.cuda.is_available(), "bench.py needs a GPU (the CUDA path has no CPU fallback)"
    torch.cuda.set_device(local_rank)
    dev = torch.device(f"cuda:{local_rank}")
    if world > 1 and not dist.is_initialized():
        dist.init_process_group("nccl", device_id=dev)
    near_cpus = bind_near_gpu(local_rank) if world > 1 else None
    peaks = load_peaks()
    cfg = CONFIGS[args.config]
    batch, k = cfg["batch"], cfg["k"]
    # The clock sampler starts BEFORE the workload is generated: with several ranks on one box, something NVML does
    # lazily about 1.5 s after a process's first queries stalled rank 0's second timed step by 70-175 ms (12-step
    # probes at 4 GPUs: present in every run with the sampler started right before warm-up, absent without the
    # sampler); started here, that happens tens of seconds before the timed region.
    sampler = ClockSampler(local_rank)
    if rank == 0 and not args.no_clock_sampler:
        sampler.start()

    if cfg["total"] is not None:  # strong scaling: this rank's contiguous share of the SAME image set
        lo, hi = shard_range(cfg["total"], rank, world)
        packed, ids, hw = make_workload(cfg["total"], seed=0, device=dev, cfg=cfg, lo=lo, hi=hi)
        n_total = cfg["total"]
    else:
        packed, ids, hw = make_workload(cfg["images_per_gpu"], seed=rank, device=dev, cfg=cfg)
        n_total = cfg["images_per_gpu"] * world
    n_local = len(packed)
    trunk = ResNet50Trunk(stage_ref.full_resnet50(seed=1234), dev, max_batch=batch)
    backend = TimedBackend(CudaBackend(trunk, lanes=2))
    stage = OutlierStage(backend, batch_size=batch, pca_components=k, class_scoring=cfg["class_scoring"])

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step(from_host=False, src=None):
        res = stage.run(src if src is not None else packed, ids, N_CLASSES, from_host=from_host)
        if from_host:  # what process_image_directory / detect_outliers hand back to the host
            out = (stage.to_host(res.features, "features"), stage.to_host(res.z, "z"),
                   stage.to_host(res.class_outliers, "class_flags"), stage.to_host(res.global_outliers, "global_flags"))
            return res, out
        # Every step ends with the host waiting for its result (the drop-in returns the flags to the caller).  Without
        # it the host runs a whole step ahead of the GPU; with 4 ranks one of the first timed steps then stalled for
        # 100-300 ms (ranks enqueue their collectives out of phase) -- with the step closed like this it does not.
        torch.cuda.current_stream().synchronize()
        return res, None

    # ---- value: inputs resident in HBM ----
    # No cyclic-GC passes inside the timed regions: a generation-2 collection over the objects the workload set-up
    # left behind cost rank 0 about 70 ms in its second timed step (every rank then waits for it at the all-reduce).
    import gc
    gc.collect()
    gc.freeze()
    gc.disable()
    for _ in range(args.warmup):
        step()
    barrier()
    backend.enabled = True
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
    clocks = sampler.stop() if rank == 0 else None
    ms_total = e0.elapsed_time(e1)
    t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_step = t.item() / args.steps
    value = n_total / (ms_step * 1e-3)

    # With two lanes the calls of one lane overlap the other lane's, so per-call durations do not add up to a time:
    # the trunk time is the length of the UNION of the calls' [start, end] intervals (device timestamps relative to
    # e0), i.e. the time during which at least one irp_resnet50_embed call was in flight -- the other lane's
    # preprocess kernels run inside those intervals too, so the figure is conservative for the convolutions.  What is
    # left of the embed phase (union of whole [preprocess start, trunk end] intervals) is the EXPOSED preprocess time.
    lanes = getattr(backend.inner, "n_lanes", 1)
    stamps = [[e0.elapsed_time(x) for x in e] for e, _ in backend.events]
    emb_ms = union_ms([(t[1], t[2]) for t in stamps])
    pre_ms = union_ms([(t[0], t[2]) for t in stamps]) - emb_ms
    # per-rank embed time (preprocess + trunk): the step ends when the SLOWEST rank reaches the PCA all-reduce, so
    # the spread between GPUs of one box shows up as waiting time in every other rank's "pca_lof_other"
    rank_embed_ms = None
    if world > 1:
        mine = torch.tensor([(pre_ms + emb_ms) / args.steps], dtype=torch.float64, device=dev)
        allr = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(allr, mine)
        rank_embed_ms = [round(float(v.item()), 3) for v in allr]
    n_emb = sum(n for _, n in backend.events)
    calls = len(backend.events)
    conv_tflops = FLOPS_PER_IMAGE * n_emb / (max(emb_ms, 1e-9) * 1e-3) / 1e12

    # ---- preprocess alone: the same batches on one stream, nothing else on the GPU (in the step its launches share
    # the SMs with the other lane's trunk kernels, so their in-step durations are not kernel times) ----
    from irp_b200 import _lib as _irp_lib, ops as _irp_ops
    pre_events = []
    for rep in range(2):  # first pass warms the allocator
        pre_events = []
        for lo in range(0, n_local, batch):
            part = packed.slice(lo, min(n_local, lo + batch))
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            x = _irp_ops.preprocess(part.pixels, part.offsets, part.hw, packed.max_taps, _irp_lib.LAYOUT_NHWC4P)
            b.record()
            pre_events.append((a, b))
            del x
        torch.cuda.synchronize()
    pre_alone_ms = sum(a.elapsed_time(b) for a, b in pre_events)
    pre_gbs = algorithmic_preprocess_bytes(hw) / (max(pre_alone_ms, 1e-9) * 1e-3) / 1e9

    # ---- one extra traced step (synchronised after every phase; outside every timed region) ----
    stage.trace = True
    stage.traces = []
    step()
    stage.trace = False
    phase_ms = dict(stage.traces[-1]) if stage.traces else None
    if phase_ms is not None and world > 1:  # the slowest rank's view of every phase
        keys = sorted(phase_ms)
        v = torch.tensor([phase_ms[q] for q in keys], dtype=torch.float64, device=dev)
        dist.all_reduce(v, op=dist.ReduceOp.MAX)
        phase_ms = {q: round(float(x), 2) for q, x in zip(keys, v.tolist())}
    barrier()

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
               "d2h_bytes_per_step": d2h, "cpus_near_gpu": near_cpus}

    gc.enable()
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    # ---- CPU baseline on a bounded sample (rank 0, N=1 only) ----
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        torch.set_num_threads(host_cores())
        sample = 1024
        images, labels = host_sample(sample, seed=0, cfg=cfg)
        sec = cpu_stage_sample(images, labels, k)
        cpu = {"value": sample / sec, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
               "sample": f"{sample} images of the same workload, once: PIL transform + torchvision ResNet-50 fp32 "
                         f"(batch 32) + sklearn PCA + LocalOutlierFactor (oracle/stage_ref.py) in {sec:.1f} s"}
    trunk_traffic, trunk_src = measured_traffic("trunk_call_batch256")
    pre_traffic, pre_src = measured_traffic("preprocess_call_batch256")
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": cfg["scaling"], "vs_baseline": None,
        "dtype": "bf16", "data": "synthetic", "config": config_dict(args.config, world),
        "roofline": {"bound": "tensor", "achieved": conv_tflops, "peak": peaks["tflops_sustained"], "unit": "TFLOP/s",
                     "frac": conv_tflops / peaks["tflops_sustained"],
                     "traffic": trunk_traffic if batch == 256 else None, "traffic_source": trunk_src,
                     "kernel": "conv_gemm2_kernel / conv_chain_kernel / conv3x3_c64_kernel / stem_pool_kernel (the 53 "
                               f"convolutions of one irp_resnet50_embed call, batch {batch})",
                     "per_launch": f"{FLOPS_PER_IMAGE:.4e} FLOP/image x {batch} images per trunk call; {calls} calls "
                                   f"on {lanes} lane(s) bracketed by CUDA events on their streams; time = union of "
                                   f"the calls' intervals = {emb_ms / max(calls, 1):.3f} ms per call",
                     "peak_source": f"{peaks['source']} bf16_tflops_sustained"},
        "roofline_preprocess": {"bound": "hbm", "achieved": pre_gbs, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                                "frac": pre_gbs / peaks["hbm_gbs"],
                                "traffic": pre_traffic if batch == 256 else None, "traffic_source": pre_src,
                                "kernel": "resample_fused_kernel (+ resample_plan_kernel)",
                                "per_launch": "3*ceil(224/232*short)^2 source bytes + 301056 output bytes per image; "
                                              f"the step's {len(pre_events)} batches timed alone with CUDA events "
                                              f"after the step loop, mean "
                                              f"{pre_alone_ms / max(len(pre_events), 1) * 1e3:.1f} us"},
        "stage_ms": {"preprocess_exposed": pre_ms / args.steps, "preprocess_alone": pre_alone_ms,
                     "trunk": emb_ms / args.steps, "lanes": lanes,
                     "pca_lof_other": ms_step - (pre_ms + emb_ms) / args.steps,
                     "embed_per_rank": rank_embed_ms, "per_step": per_step_ms,
                     "phases_of_one_traced_step": phase_ms},
        "cpu_baseline": cpu, "e2e": e2e, "clocks": clocks,
        "gpu_launches": launches_per_step(n_local, batch, 2048, k, cfg["class_scoring"]) * args.steps,
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
    ap.add_argument("--config", choices=sorted(CONFIGS), default="cfg2")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-clock-sampler", action="store_true")
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
