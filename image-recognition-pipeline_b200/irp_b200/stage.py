"""Host-side engine of the B200 outlier-detection stage.

``ResNet50Trunk``   owns an ``irp_resnet50`` handle: loads a torchvision ResNet-50 state (BN folded on the device)
                    and embeds padded NHWC4 bf16 batches.
``OutlierStage``    runs the whole hot path on one rank's shard of a packed uint8 image batch:
                    preprocess -> embed -> (all-reduce of the PCA partial sums) -> PCA fit -> project ->
                    LOF per class + global, mirroring functions/data_curation.py:661-728 of the reference.

Multi-GPU (one process per GPU, torch.distributed/NCCL): images shard across ranks with no data-path collective;
the PCA needs ONE all-reduce of [count, sum, scatter] (plus an 8 KB broadcast of the shift vector so every rank
accumulates about the same origin); scoring needs the projected rows of all ranks (one all-gather of N x k
floats), after which each rank runs the neighbour search for its share of the query tiles and three length-N fp64
vectors (k-distance, lrd, score) are summed across ranks.
"""
from __future__ import annotations

import ctypes as C
import math
import os
from dataclasses import dataclass
from typing import List, Optional, Sequence

import numpy as np
import torch

from . import _lib, ops


def conv_bn_pairs(model: torch.nn.Module):
    """(conv, bn) pairs of a torchvision ResNet-50 in the library's index order:
    conv1, then per bottleneck conv1, conv2, conv3[, downsample] (irp_resnet50_conv_shape)."""
    pairs = [(model.conv1, model.bn1)]
    for layer in (model.layer1, model.layer2, model.layer3, model.layer4):
        for blk in layer:
            pairs += [(blk.conv1, blk.bn1), (blk.conv2, blk.bn2), (blk.conv3, blk.bn3)]
            if blk.downsample is not None:
                pairs.append((blk.downsample[0], blk.downsample[1]))
    if len(pairs) != _lib.NUM_CONVS:
        raise ValueError(f"expected a ResNet-50 ({_lib.NUM_CONVS} convolutions), found {len(pairs)}")
    return pairs


class ResNet50Trunk:
    """ResNet-50 minus fc on the library's tcgen05 kernels (functions/data_curation.py:654-659, :677)."""

    def __init__(self, torch_model: torch.nn.Module, device: torch.device, max_batch: int = 256):
        self._model = torch_model  # kept so that lane() can build a second handle from the same weights
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("ResNet50Trunk needs a CUDA (sm_100) device; there is no CPU path")
        idx = self.device.index if self.device.index is not None else torch.cuda.current_device()
        self.lib = _lib.init(idx)
        self.max_batch = int(max_batch)
        self._handle = C.c_void_p()
        with torch.cuda.device(idx):
            _lib.check(self.lib.irp_resnet50_create(C.byref(self._handle), self.max_batch), "irp_resnet50_create")
            stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
            for i, (conv, bn) in enumerate(conv_bn_pairs(torch_model)):
                shape = [C.c_int() for _ in range(5)]
                _lib.check(self.lib.irp_resnet50_conv_shape(i, *[C.byref(s) for s in shape]), "conv_shape")
                cout, cin, kh, kw, _ = (s.value for s in shape)
                w = conv.weight.detach().to(self.device, torch.float32).contiguous()
                if tuple(w.shape) != (cout, cin, kh, kw):
                    raise ValueError(f"conv {i}: weight shape {tuple(w.shape)} != {(cout, cin, kh, kw)}")
                if conv.bias is not None:
                    raise ValueError("ResNet-50 convolutions are bias-free (torchvision/models/resnet.py)")
                t = [p.detach().to(self.device, torch.float32).contiguous()
                     for p in (bn.weight, bn.bias, bn.running_mean, bn.running_var)]
                _lib.check(self.lib.irp_resnet50_load_conv(self._handle, i, C.c_void_p(w.data_ptr()),
                                                           *[C.c_void_p(p.data_ptr()) for p in t],
                                                           C.c_float(bn.eps), stream), f"load_conv[{i}]")
            torch.cuda.synchronize(idx)

    def lanes(self, n: int) -> List["ResNet50Trunk"]:
        """[self] + (n - 1) further handles (own activation arenas and tensor maps) built from the same weights, for
        CudaBackend(lanes=n); created once and shared by every stage that runs on this trunk, closed with it."""
        extra = self.__dict__.setdefault("_lane_handles", [])
        while len(extra) < n - 1:
            extra.append(ResNet50Trunk(self._model, self.device, self.max_batch))
        return [self] + extra[: n - 1]

    @property
    def handle(self) -> int:
        return int(self._handle.value)

    def embed(self, x_nhwc4p: torch.Tensor) -> torch.Tensor:
        """bf16 [B,230,230,4] -> fp32 [B,2048]; B <= max_batch."""
        if x_nhwc4p.shape[0] > self.max_batch:
            raise ValueError(f"batch {x_nhwc4p.shape[0]} > max_batch {self.max_batch}")
        return ops.resnet50_embed(self.handle, x_nhwc4p)

    def embed_nchw(self, x: torch.Tensor) -> torch.Tensor:
        """Normalised [B,3,224,224] tensor (what the reference feeds `model`) -> fp32 [B,2048]."""
        b = x.shape[0]
        xp = torch.zeros((b, _lib.PAD_HW, _lib.PAD_HW, 4), dtype=torch.bfloat16, device=self.device)
        xp[:, 3:3 + _lib.CROP, 3:3 + _lib.CROP, :3] = x.to(self.device).permute(0, 2, 3, 1).to(torch.bfloat16)
        out = []
        for s in range(0, b, self.max_batch):
            out.append(self.embed(xp[s:s + self.max_batch].contiguous()))
        return torch.cat(out, 0)

    def close(self):
        for t in self.__dict__.pop("_lane_handles", []):
            t.close()
        if self._handle:
            self.lib.irp_resnet50_destroy(self._handle)
            self._handle = C.c_void_p()

    def __del__(self):  # pragma: no cover
        try:
            self.close()
        except Exception:
            pass


# -----------------------------------------------------------------------------------------------------------------
# packing a ragged batch of decoded images
# -----------------------------------------------------------------------------------------------------------------
ALIGN = 128  # every image starts on a 128-byte boundary of the packed buffer


@dataclass
class PackedImages:
    """A ragged uint8 HWC batch packed back to back.  `pixels/offsets/hw` are torch tensors (host or device);
    `offsets_np/hw_np` are always-host copies used for slicing without a device round trip."""
    pixels: torch.Tensor   # uint8 [total_bytes]
    offsets: torch.Tensor  # int64 [n]
    hw: torch.Tensor       # int32 [n,2]
    max_taps: int
    offsets_np: np.ndarray = None
    hw_np: np.ndarray = None

    def __post_init__(self):
        if self.offsets_np is None:
            self.offsets_np = self.offsets.cpu().numpy()
        if self.hw_np is None:
            self.hw_np = self.hw.cpu().numpy()

    def __len__(self):
        return int(self.hw_np.shape[0])

    def to(self, device, non_blocking=True):
        return PackedImages(self.pixels.to(device, non_blocking=non_blocking),
                            self.offsets.to(device, non_blocking=non_blocking),
                            self.hw.to(device, non_blocking=non_blocking), self.max_taps, self.offsets_np, self.hw_np)

    def slice(self, lo: int, hi: int) -> "PackedImages":
        """Images [lo, hi) as a view (offsets rebased)."""
        start = int(self.offsets_np[lo])
        h, w = int(self.hw_np[hi - 1, 0]), int(self.hw_np[hi - 1, 1])
        end = int(self.offsets_np[hi - 1]) + h * w * 3
        return PackedImages(self.pixels[start:end], self.offsets[lo:hi] - start, self.hw[lo:hi], self.max_taps,
                            self.offsets_np[lo:hi] - start, self.hw_np[lo:hi])

    def nbytes(self) -> int:
        return int((self.hw_np[:, 0].astype(np.int64) * self.hw_np[:, 1] * 3).sum())


def taps_for(h: int, w: int, transform: int = 0) -> int:
    """2*ceil(max(scale,1))+1 for the resize transform of an h x w image (matches irp_preprocess_geometry_ex):
    transform 0 = short side -> 232, 1 = Resize((256,256)) of the classifier's val_transform."""
    if transform == _lib.TRANSFORM_VAL_256:
        return 2 * int(math.ceil(max(h / 256, w / 256, 1.0))) + 1
    if transform in (_lib.TRANSFORM_WDS_LANCZOS, _lib.TRANSFORM_HASH_64):
        return _lib.geometry(h, w, transform)[4]
    if w <= h:
        out_w, out_h = 232, int(232 * h / w)
    else:
        out_h, out_w = 232, int(232 * w / h)
    s = max(h / out_h, w / out_w, 1.0)
    return 2 * int(math.ceil(s)) + 1


def pack_images(images: Sequence[np.ndarray], pin: bool = True, transform: int = 0) -> PackedImages:
    """Pack HWC uint8 RGB arrays into one (pinned) host buffer; `transform` only sizes the tap bound."""
    sizes = np.array([[im.shape[0], im.shape[1]] for im in images], np.int32).reshape(-1, 2)
    nbytes = sizes[:, 0].astype(np.int64) * sizes[:, 1] * 3
    padded = (nbytes + ALIGN - 1) // ALIGN * ALIGN
    offsets = np.concatenate([[0], np.cumsum(padded)[:-1]]).astype(np.int64)
    total = int(padded.sum())
    buf = torch.empty(max(total, ALIGN), dtype=torch.uint8, pin_memory=pin and torch.cuda.is_available())
    view = buf.numpy()
    for im, o, nb in zip(images, offsets, nbytes):
        if im.dtype != np.uint8 or im.ndim != 3 or im.shape[2] != 3:
            raise ValueError(f"expected HWC uint8 RGB, got {im.dtype} {im.shape}")
        view[o:o + nb] = np.ascontiguousarray(im).reshape(-1)
    taps = max((taps_for(int(h), int(w), transform) for h, w in sizes), default=3)
    return PackedImages(buf[:max(total, 1)], torch.from_numpy(offsets), torch.from_numpy(sizes), taps, offsets, sizes)


# -----------------------------------------------------------------------------------------------------------------
# the stage
# -----------------------------------------------------------------------------------------------------------------
@dataclass
class PCAState:
    mean: torch.Tensor         # fp64 [d]
    components: torch.Tensor   # fp64 [k,d]
    explained_variance: torch.Tensor  # fp64 [k]
    total_variance: float
    n_samples: int


@dataclass
class StageResult:
    features: torch.Tensor      # fp32 [n_local, 2048]
    z: torch.Tensor             # fp32 [n_total, k] (all ranks' rows, rank order)
    pca: PCAState
    class_outliers: torch.Tensor   # bool [n_total]
    global_outliers: torch.Tensor  # bool [n_total]
    class_scores: torch.Tensor     # fp64 [n_total]
    global_scores: torch.Tensor    # fp64 [n_total]


class CudaBackend:
    """The product backend: every step is a libirp_b200 kernel (through the irp_b200 custom ops)."""

    def __init__(self, trunk: ResNet50Trunk, lanes: int = 1):
        """lanes > 1: batches alternate between `lanes` trunk handles, each fed on its own stream, so that the tail of
        one lane's kernel (the last, partly filled wave of a persistent grid; the pipeline fill of the next one) is
        covered by the other lane's launches: + 3 % on the embed phase with two lanes, nothing more with three
        (tools/lane_probe.py; limiting each lane to half of the SMs instead was measured SLOWER than one lane)."""
        self.trunk = trunk
        self.device = trunk.device
        self.n_lanes = max(1, int(lanes))
        self.lane_trunks: List[ResNet50Trunk] = [trunk]
        self.lane_streams: List[torch.cuda.Stream] = []

    def ensure_lanes(self):
        """The extra handles (activation arenas) and streams are created at the first multi-batch pass."""
        if self.n_lanes > 1 and not self.lane_streams:
            self.lane_trunks = self.trunk.lanes(self.n_lanes)
            self.lane_streams = [torch.cuda.Stream(device=self.device) for _ in range(self.n_lanes)]

    def embed(self, part: PackedImages, max_taps: int, lane: Optional[int] = None) -> torch.Tensor:
        x = ops.preprocess(part.pixels, part.offsets, part.hw, max_taps, _lib.LAYOUT_NHWC4P)
        trunk = self.trunk if lane is None else self.lane_trunks[lane]
        return trunk.embed(x)

    cov_accumulate = staticmethod(ops.cov_accumulate)
    pca_fit = staticmethod(ops.pca_fit)
    pca_transform = staticmethod(ops.pca_transform)
    lof = staticmethod(ops.lof)
    lof_sharded = staticmethod(ops.lof_sharded)
    lof_sharded_multi = staticmethod(ops.lof_sharded_multi)


class OutlierStage:
    """embed + PCA + outlier scoring for one rank's shard (see module docstring).

    `backend` supplies the compute steps (CudaBackend in the product; the CPU test-suite injects an oracle-backed
    one to exercise the sharding / all-reduce / gather logic under gloo)."""

    def __init__(self, backend, batch_size: int = 256, pca_components: int = 50, class_n_neighbors: int = 30,
                 class_contamination: float = 0.05, global_n_neighbors: int = 75,
                 global_contamination: float = 0.03, process_group=None, embed_dim: int = _lib.EMBED_DIM,
                 class_scoring: bool = True, trace: bool = False, lanes: int = 2):
        if isinstance(backend, ResNet50Trunk):
            backend = CudaBackend(backend, lanes=lanes)  # the second lane's arena is allocated at its first use
        self.backend = backend
        self.device = torch.device(backend.device)
        self.batch_size = int(batch_size)
        if hasattr(backend, "trunk"):
            self.batch_size = min(self.batch_size, backend.trunk.max_batch)
        self.k = int(pca_components)
        self.embed_dim = int(embed_dim)
        self.class_nn, self.class_cont = int(class_n_neighbors), float(class_contamination)
        self.global_nn, self.global_cont = int(global_n_neighbors), float(global_contamination)
        self.pg = process_group
        self.class_scoring = bool(class_scoring)  # False: global scorer only (BASELINE configs[3])
        self.trace = bool(trace)                  # per-phase wall-clock times in self.traces (synchronises)
        self.copy_stream = torch.cuda.Stream(device=self.device) if self.device.type == "cuda" else None
        self._pinned = {}

    # ---- device -> host ----
    def to_host(self, t: torch.Tensor, slot: str = "default") -> torch.Tensor:
        """Copy a device tensor into a reusable PINNED host buffer (one per `slot`) and return the host view.

        `tensor.cpu()` lands in pageable memory and runs at ~2 GB/s (100 ms for the 221 MB feature matrix of 27k
        images); a pinned destination makes the same copy a single ~5 ms DMA.  The view is valid until the next call
        with the same slot."""
        if not t.is_cuda:
            return t
        n = t.numel() * t.element_size()
        buf = self._pinned.get(slot)
        if buf is None or buf.numel() < n:
            buf = torch.empty(max(n, 1), dtype=torch.uint8, pin_memory=True)
            self._pinned[slot] = buf
        host = buf[:n].view(t.dtype).view(t.shape)
        host.copy_(t.contiguous(), non_blocking=True)
        torch.cuda.current_stream(t.device).synchronize()
        return host

    # ---- distributed helpers ----
    def _dist(self):
        import torch.distributed as dist
        return dist if (dist.is_available() and dist.is_initialized()) else None

    @property
    def world_size(self) -> int:
        d = self._dist()
        return d.get_world_size(self.pg) if d else 1

    # ---- embed ----
    def embed_packed(self, packed: PackedImages, from_host: bool = False) -> torch.Tensor:
        """Preprocess + embed every image of `packed`; fp32 [n,2048] on the device.

        Batch j runs on lane j % lanes (CudaBackend(lanes=...): one trunk handle and one stream per lane; one lane =
        the caller's stream); the caller's stream waits for every lane at the end.  With from_host=True the packed
        tensors live in (pinned) host memory: every batch is copied on the copy stream into one slot of a small ring
        of device staging buffers that is reused across batches and calls (no allocation inside the loop), the copy
        of a batch overlapping the compute of the batches before it."""
        n = len(packed)
        feats = torch.empty((n, self.embed_dim), dtype=torch.float32, device=self.device)
        if n == 0:
            return feats
        bs = self.batch_size
        bounds = list(range(0, n, bs)) + [n]
        nb = len(bounds) - 1
        if self.device.type != "cuda":  # oracle-backed test backend
            for j in range(nb):
                part = packed.slice(bounds[j], bounds[j + 1])
                feats[bounds[j]:bounds[j + 1]] = self.backend.embed(part, packed.max_taps)
            return feats

        lanes = getattr(self.backend, "n_lanes", 1) if nb > 1 else 1
        main = torch.cuda.current_stream(self.device)
        if lanes > 1:
            self.backend.ensure_lanes()
            streams = self.backend.lane_streams
            fork = torch.cuda.Event()
            fork.record(main)
            for s in streams:
                s.wait_event(fork)
        else:
            streams = [main]
        ring = self._h2d_ring(packed, bounds, lanes + 1) if from_host else None
        if ring is not None and lanes == 1:
            fork = torch.cuda.Event()
            fork.record(main)
        if ring is not None:
            self.copy_stream.wait_event(fork)  # staging slots last used by an earlier call on this stream

        for j in range(nb):
            lo, hi = bounds[j], bounds[j + 1]
            lane = j % lanes
            s = streams[lane]
            if ring is not None:
                part, slot = self._stage_h2d(ring, j, packed, lo, hi)
            with torch.cuda.stream(s):
                if ring is not None:
                    s.wait_event(slot["ready"])
                else:
                    # device-resident input: the slice's rebased offsets are computed ON the lane's stream
                    part = packed.slice(lo, hi)
                out = self.backend.embed(part, packed.max_taps, lane=lane) if lanes > 1 else \
                    self.backend.embed(part, packed.max_taps)
                feats[lo:hi] = out
                if ring is not None:
                    slot["done"] = torch.cuda.Event()
                    slot["done"].record(s)
        if lanes > 1:
            for s in streams:
                done = torch.cuda.Event()
                done.record(s)
                main.wait_event(done)
        return feats

    def _h2d_ring(self, packed: PackedImages, bounds, slots: int):
        """Device staging buffers for host-resident input: `slots` x (pixels, offsets, hw), sized for the largest
        batch; kept across calls and grown on demand."""
        nbytes = 0
        for j in range(len(bounds) - 1):
            lo, hi = bounds[j], bounds[j + 1]
            h, w = int(packed.hw_np[hi - 1, 0]), int(packed.hw_np[hi - 1, 1])
            nbytes = max(nbytes, int(packed.offsets_np[hi - 1]) + h * w * 3 - int(packed.offsets_np[lo]))
        rows = max(bounds[j + 1] - bounds[j] for j in range(len(bounds) - 1))
        ring = getattr(self, "_ring", None)
        if ring is None or len(ring) < slots or ring[0]["pixels"].numel() < nbytes or ring[0]["hw"].shape[0] < rows:
            cap = (max(nbytes, 1) * 9 // 8 + 255) // 256 * 256  # headroom: the next call's batches differ a little
            ring = [{"pixels": torch.empty(cap, dtype=torch.uint8, device=self.device),
                     "offsets": torch.empty(rows, dtype=torch.int64, device=self.device),
                     "hw": torch.empty((rows, 2), dtype=torch.int32, device=self.device),
                     "ready": None, "done": None} for _ in range(slots)]
            self._ring = ring
        return ring

    def _stage_h2d(self, ring, j: int, packed: PackedImages, lo: int, hi: int):
        """Copy images [lo, hi) of the host-resident `packed` into ring slot j % len(ring) on the copy stream."""
        slot = ring[j % len(ring)]
        start = int(packed.offsets_np[lo])
        h, w = int(packed.hw_np[hi - 1, 0]), int(packed.hw_np[hi - 1, 1])
        end = int(packed.offsets_np[hi - 1]) + h * w * 3
        m = hi - lo
        cs = self.copy_stream
        with torch.cuda.stream(cs):
            if slot["done"] is not None:
                cs.wait_event(slot["done"])  # the batch that used this slot has been consumed
            px = slot["pixels"][: end - start]
            off = slot["offsets"][:m]
            hw = slot["hw"][:m]
            px.copy_(packed.pixels[start:end], non_blocking=True)
            off.copy_(packed.offsets[lo:hi], non_blocking=True)
            off.sub_(start)  # rebased on the device: no pageable temporary, no blocking copy
            hw.copy_(packed.hw[lo:hi], non_blocking=True)
            slot["ready"] = torch.cuda.Event()
            slot["ready"].record(cs)
        part = PackedImages(px, off, hw, packed.max_taps, packed.offsets_np[lo:hi] - start, packed.hw_np[lo:hi])
        return part, slot

    # ---- PCA ----
    def fit_pca(self, feats: torch.Tensor) -> PCAState:
        d = feats.shape[1]
        dist = self._dist()
        multi = dist is not None and self.world_size > 1
        # common origin for every rank: the mean of rank 0's first rows
        shift = feats[: min(256, feats.shape[0])].mean(0).contiguous() if feats.shape[0] > 0 else \
            torch.zeros(d, dtype=torch.float32, device=self.device)
        if multi:
            src = dist.get_global_rank(self.pg, 0) if self.pg is not None else 0
            dist.broadcast(shift, src=src, group=self.pg)
        # one flat fp64 buffer [count | sum | scatter] so the reduction is a single all-reduce
        acc = torch.zeros(1 + d + d * d, dtype=torch.float64, device=self.device)
        count, total, scatter = acc[:1], acc[1:1 + d], acc[1 + d:].view(d, d)
        if feats.shape[0] > 0:
            self.backend.cov_accumulate(feats, shift, count, total, scatter)
        if multi:
            dist.all_reduce(acc, op=dist.ReduceOp.SUM, group=self.pg)
        mean, comps, evals = self.backend.pca_fit(count, total, scatter, shift, self.k)
        n_total = int(round(float(count.item())))
        return PCAState(mean, comps, evals[: self.k].clone(), float(evals[self.k].item()), n_total)

    def transform(self, feats: torch.Tensor, pca: PCAState) -> torch.Tensor:
        if feats.shape[0] == 0:
            return torch.empty((0, self.k), dtype=torch.float32, device=self.device)
        return self.backend.pca_transform(feats, pca.mean, pca.components)

    # ---- scoring ----
    def gather_rows(self, z_local: torch.Tensor, ids_local: torch.Tensor):
        """All ranks' projected rows and class ids, concatenated in rank order (identity on one rank).

        ONE all-gather of the row counts and ONE of a padded [rows, k + 1] fp32 block (ids ride along as an exactly
        representable float column), instead of separate exchanges for z and ids."""
        dist = self._dist()
        if not dist or self.world_size == 1:
            return z_local, ids_local
        ws = self.world_size
        n_local = torch.tensor([z_local.shape[0]], dtype=torch.int64, device=self.device)
        sizes_t = torch.empty(ws, dtype=torch.int64, device=self.device)
        dist.all_gather_into_tensor(sizes_t, n_local, group=self.pg)
        sizes = [int(v) for v in sizes_t.cpu().tolist()]
        mx = max(sizes)
        k = z_local.shape[1]
        pad = torch.zeros((mx, k + 1), dtype=torch.float32, device=self.device)
        pad[: z_local.shape[0], :k] = z_local
        pad[: z_local.shape[0], k] = ids_local.to(torch.float32)  # class ids < 2^24: exact
        out = torch.empty((ws * mx, k + 1), dtype=torch.float32, device=self.device)
        dist.all_gather_into_tensor(out, pad, group=self.pg)
        if all(sz == mx for sz in sizes):
            rows = out
        else:
            rows = torch.cat([out[r * mx: r * mx + sizes[r]] for r in range(ws)], 0)
        return rows[:, :k].contiguous(), rows[:, k].to(torch.int32).contiguous()

    def detect(self, z_all: torch.Tensor, class_ids_all: torch.Tensor, n_classes: int):
        """Per-class + global LOF over ALL rows.  On several ranks the O(n^2) neighbour search is sharded: rank r
        searches the query tiles r, r+W, r+2W, ... and three [problems, n] fp64 buffers are summed across ranks (one
        all-reduce per LOF phase for both scorers together)."""
        dist = self._dist()
        problems = []
        if self.class_scoring:
            problems.append((class_ids_all, n_classes, self.class_nn, self.class_cont))
        problems.append((None, 1, self.global_nn, self.global_cont))
        if dist is not None and self.world_size > 1 and hasattr(self.backend, "lof_sharded_multi"):
            rank, ws = dist.get_rank(self.pg), self.world_size

            def all_reduce(t):
                dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.pg)

            res = self.backend.lof_sharded_multi(z_all, problems, rank, ws, all_reduce)
        elif len(problems) == 2 and self.device.type == "cuda":
            # one rank: the per-class and the global problem are independent -> the per-class one runs on a side
            # stream beside the global one (its grid leaves SMs idle in the last wave, and its group-id validation
            # synchronises only the side stream)
            main = torch.cuda.current_stream(self.device)
            if getattr(self, "_lof_stream", None) is None:
                self._lof_stream = torch.cuda.Stream(device=self.device)
            side = self._lof_stream
            fork = torch.cuda.Event()
            fork.record(main)
            side.wait_event(fork)
            g_res = self.backend.lof(z_all, *problems[1][:1], *problems[1][1:])
            with torch.cuda.stream(side):
                c_res = self.backend.lof(z_all, *problems[0][:1], *problems[0][1:])
                done = torch.cuda.Event()
                done.record(side)
            main.wait_event(done)
            for t in c_res:
                t.record_stream(main)
            res = [c_res, g_res]
        else:
            res = [self.backend.lof(z_all, g, ng, k, c) for (g, ng, k, c) in problems]
        gs, _, gf = res[-1]
        if self.class_scoring:
            cs, _, cf = res[0]
        else:
            cs, cf = torch.zeros_like(gs), torch.zeros_like(gf)
        return cf.bool(), gf.bool(), cs, gs

    # ---- whole stage ----
    def run(self, packed: PackedImages, class_ids: torch.Tensor, n_classes: int,
            from_host: bool = False) -> StageResult:
        """`packed` / `class_ids` are THIS rank's shard; returned flags cover all ranks' rows in rank order."""
        trace = self._trace_begin()
        feats = self.embed_packed(packed, from_host=from_host)
        self._trace(trace, "embed")
        pca = self.fit_pca(feats)
        self._trace(trace, "fit_pca")
        z_local = self.transform(feats, pca)
        ids_local = class_ids.to(self.device, non_blocking=True).to(torch.int32)
        self._trace(trace, "transform")
        z_all, ids_all = self.gather_rows(z_local, ids_local)
        self._trace(trace, "gather")
        cf, gf, cs, gs = self.detect(z_all.contiguous(), ids_all.contiguous(), n_classes)
        self._trace(trace, "detect")
        return StageResult(feats, z_all, pca, cf, gf, cs, gs)

    # ---- optional per-phase wall-clock trace (OutlierStage(trace=True); synchronises after every phase) ----
    def _trace_begin(self):
        if not self.trace:
            return None
        import time
        torch.cuda.synchronize(self.device)
        return [("start", time.perf_counter())]

    def _trace(self, trace, label):
        if trace is None:
            return
        import time
        torch.cuda.synchronize(self.device)
        trace.append((label, time.perf_counter()))
        if label == "detect":
            if not hasattr(self, "traces"):
                self.traces = []
            self.traces.append([(b[0], round(1e3 * (b[1] - a[1]), 2)) for a, b in zip(trace[:-1], trace[1:])])


def shard_range(n: int, rank: int, world_size: int):
    """Contiguous [lo, hi) image range of `rank` (contiguous N/G shards)."""
    base, rem = divmod(n, world_size)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)
