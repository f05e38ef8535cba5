"""torch custom ops (namespace ``irp_b200``) over the C ABI of libirp_b200.so.

Each op takes CUDA tensors, enqueues the library call on the current torch stream and returns torch tensors;
torch is plumbing here (device memory, streams), every kernel is the library's own sm_100a code.  There is no
CPU implementation behind these ops: they raise on non-CUDA tensors.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Tuple

import torch

from . import _lib


def _ptr(t: Optional[torch.Tensor]) -> C.c_void_p:
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def _stream(t: torch.Tensor) -> C.c_void_p:
    """The current torch stream OF THE TENSOR'S DEVICE (not of the process-wide current device)."""
    return C.c_void_p(torch.cuda.current_stream(t.device).cuda_stream)


def _lib_for(t: torch.Tensor):
    if not t.is_cuda:
        raise RuntimeError("irp_b200 ops run on CUDA tensors only (no CPU fallback on the outlier-stage hot path)")
    return _lib.init(t.device.index if t.device.index is not None else torch.cuda.current_device())


def _on_device_of(arg: int = 0):
    """Run the op with the device of its `arg`-th tensor argument current: the library launches on, and sets kernel
    attributes for, the CUDA runtime's current device, which need not be the tensors' device in a multi-GPU process."""
    def deco(fn):
        import functools

        @functools.wraps(fn)
        def wrapped(*a, **kw):
            t = a[arg]
            if isinstance(t, torch.Tensor) and t.is_cuda:
                with torch.cuda.device(t.device):
                    return fn(*a, **kw)
            return fn(*a, **kw)
        return wrapped
    return deco


# ---------------------------------------------------------------------------------------------------------------
# A1 preprocess
# ---------------------------------------------------------------------------------------------------------------
@torch.library.custom_op("irp_b200::preprocess", mutates_args=())
@_on_device_of(0)
def preprocess(pixels: torch.Tensor, offsets: torch.Tensor, hw: torch.Tensor, max_taps: int,
               layout: int) -> torch.Tensor:
    """Ragged uint8 HWC batch -> bf16 [n,3,224,224] (layout 0) or padded NHWC4 [n,230,230,4] (layout 1)."""
    lib = _lib_for(pixels)
    assert pixels.dtype == torch.uint8 and offsets.dtype == torch.int64 and hw.dtype == torch.int32
    n = hw.shape[0]
    shape = (n, 3, _lib.CROP, _lib.CROP) if layout == _lib.LAYOUT_NCHW else (n, _lib.PAD_HW, _lib.PAD_HW, 4)
    out = torch.empty(shape, dtype=torch.bfloat16, device=pixels.device)
    ws_bytes = lib.irp_preprocess_workspace_bytes(n, max_taps)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=pixels.device)
    _lib.check(lib.irp_preprocess(_ptr(pixels), _ptr(offsets), _ptr(hw), n, max_taps, _ptr(ws), ws_bytes, _ptr(out),
                                  layout, _stream(pixels)), "irp_preprocess")
    return out


@preprocess.register_fake
def _(pixels, offsets, hw, max_taps, layout):
    n = hw.shape[0]
    shape = (n, 3, _lib.CROP, _lib.CROP) if layout == _lib.LAYOUT_NCHW else (n, _lib.PAD_HW, _lib.PAD_HW, 4)
    return pixels.new_empty(shape, dtype=torch.bfloat16)


def _preprocess_out(n: int, layout: int, transform: int = 0):
    if layout == _lib.LAYOUT_U8_HWC:
        side = _lib.HASH_SIZE if transform == _lib.TRANSFORM_HASH_64 else _lib.CROP
        return (n, side, side, 3), torch.uint8
    if layout == _lib.LAYOUT_NCHW:
        return (n, 3, _lib.CROP, _lib.CROP), torch.bfloat16
    return (n, _lib.PAD_HW, _lib.PAD_HW, 4), torch.bfloat16


@torch.library.custom_op("irp_b200::preprocess_ex", mutates_args=())
@_on_device_of(0)
def preprocess_ex(pixels: torch.Tensor, offsets: torch.Tensor, hw: torch.Tensor, max_taps: int, layout: int,
                  transform: int) -> torch.Tensor:
    """`preprocess` with the resize geometry / filter selected by `transform` (_lib.TRANSFORM_*): 1 = the
    classifier's validation transform (functions/dataload.py:51-56), 2 = the WebDataset stage's Lanczos
    resize_and_crop_image (functions/data_curation.py:883-913); layout 2 returns the uint8 pixels [n,224,224,3]."""
    lib = _lib_for(pixels)
    assert pixels.dtype == torch.uint8 and offsets.dtype == torch.int64 and hw.dtype == torch.int32
    n = hw.shape[0]
    shape, dtype = _preprocess_out(n, layout, transform)
    out = torch.empty(shape, dtype=dtype, device=pixels.device)
    ws_bytes = lib.irp_preprocess_workspace_bytes(n, max_taps)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=pixels.device)
    _lib.check(lib.irp_preprocess_ex(_ptr(pixels), _ptr(offsets), _ptr(hw), n, max_taps, _ptr(ws), ws_bytes,
                                     _ptr(out), layout, transform, _stream(pixels)), "irp_preprocess_ex")
    return out


@preprocess_ex.register_fake
def _(pixels, offsets, hw, max_taps, layout, transform):
    shape, dtype = _preprocess_out(hw.shape[0], layout, transform)
    return pixels.new_empty(shape, dtype=dtype)


@_on_device_of(0)
def md5_rows(data: torch.Tensor) -> torch.Tensor:
    """RFC 1321 MD5 of every row of a contiguous uint8 [n, row_bytes] CUDA tensor -> uint8 [n, 16] digests."""
    lib = _lib_for(data)
    assert data.dtype == torch.uint8 and data.is_contiguous() and data.dim() == 2
    n, row_bytes = data.shape
    digest = torch.empty((n, 16), dtype=torch.uint8, device=data.device)
    _lib.check(lib.irp_md5_rows(_ptr(data), n, row_bytes, _ptr(digest), _stream(data)), "irp_md5_rows")
    return digest


def image_hashes(pixels: torch.Tensor, offsets: torch.Tensor, hw: torch.Tensor, max_taps: int) -> torch.Tensor:
    """compute_image_hash (functions/data_curation.py:283-292) of a packed batch of RGB images: Pillow-exact bicubic
    resize to 64x64 and md5 of the 12 288 bytes, both on the device -> uint8 [n, 16]."""
    small = preprocess_ex(pixels, offsets, hw, max_taps, _lib.LAYOUT_U8_HWC, _lib.TRANSFORM_HASH_64)
    return md5_rows(small.view(small.shape[0], -1))


# ---------------------------------------------------------------------------------------------------------------
# A2 ResNet-50 trunk (handle passed as an integer address)
# ---------------------------------------------------------------------------------------------------------------
@torch.library.custom_op("irp_b200::resnet50_embed", mutates_args=())
@_on_device_of(1)
def resnet50_embed(handle: int, x_nhwc4p: torch.Tensor) -> torch.Tensor:
    """bf16 [B,230,230,4] -> fp32 [B,2048] pooled embeddings."""
    lib = _lib_for(x_nhwc4p)
    assert x_nhwc4p.dtype == torch.bfloat16 and x_nhwc4p.is_contiguous()
    b = x_nhwc4p.shape[0]
    out = torch.empty((b, _lib.EMBED_DIM), dtype=torch.float32, device=x_nhwc4p.device)
    _lib.check(lib.irp_resnet50_embed(C.c_void_p(handle), _ptr(x_nhwc4p), b, _ptr(out), _stream(x_nhwc4p)),
               "irp_resnet50_embed")
    return out


@resnet50_embed.register_fake
def _(handle, x_nhwc4p):
    return x_nhwc4p.new_empty((x_nhwc4p.shape[0], _lib.EMBED_DIM), dtype=torch.float32)


# ---------------------------------------------------------------------------------------------------------------
# A3 PCA
# ---------------------------------------------------------------------------------------------------------------
@torch.library.custom_op("irp_b200::cov_accumulate", mutates_args=("count", "total", "scatter"))
@_on_device_of(0)
def cov_accumulate(x: torch.Tensor, shift: torch.Tensor, count: torch.Tensor, total: torch.Tensor,
                   scatter: torch.Tensor) -> None:
    """count += n; total += sum(x - shift); scatter += (x - shift)^T (x - shift)  (fp64 accumulators)."""
    lib = _lib_for(x)
    assert x.dtype == torch.float32 and x.is_contiguous() and shift.dtype == torch.float32
    assert count.dtype == total.dtype == scatter.dtype == torch.float64
    n, d = x.shape
    ws_bytes = lib.irp_cov_workspace_bytes(n, d)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=x.device)
    _lib.check(lib.irp_cov_accumulate(_ptr(x), n, d, _ptr(shift), _ptr(count), _ptr(total), _ptr(scatter), _ptr(ws),
                                      ws_bytes, _stream(x)), "irp_cov_accumulate")


@torch.library.custom_op("irp_b200::pca_fit", mutates_args=())
@_on_device_of(0)
def pca_fit(count: torch.Tensor, total: torch.Tensor, scatter: torch.Tensor, shift: torch.Tensor,
            k: int) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """-> (mean fp64[d], components fp64[k,d], eigenvalues fp64[k+1] = top-k then total variance)."""
    lib = _lib_for(scatter)
    d = scatter.shape[0]
    mean = torch.empty(d, dtype=torch.float64, device=scatter.device)
    comps = torch.empty((k, d), dtype=torch.float64, device=scatter.device)
    evals = torch.empty(k + 1, dtype=torch.float64, device=scatter.device)
    ws_bytes = lib.irp_pca_fit_workspace_bytes(d, k)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=scatter.device)
    _lib.check(lib.irp_pca_fit(_ptr(count), _ptr(total), _ptr(scatter), _ptr(shift), d, k, _ptr(mean), _ptr(comps),
                               _ptr(evals), _ptr(ws), ws_bytes, _stream(count)), "irp_pca_fit")
    return mean, comps, evals


@_on_device_of(0)
def pca_fit_ex(count: torch.Tensor, total: torch.Tensor, scatter: torch.Tensor, shift: torch.Tensor, k: int,
               solver: int = 0, lanczos_first_check: int = 0):
    """pca_fit with the eigensolver forced (_lib.PCA_SOLVER_*) and the library's report:
    -> (mean, components, eigenvalues, {"solver", "lanczos_steps", "checks"})."""
    lib = _lib_for(scatter)
    d = scatter.shape[0]
    mean = torch.empty(d, dtype=torch.float64, device=scatter.device)
    comps = torch.empty((k, d), dtype=torch.float64, device=scatter.device)
    evals = torch.empty(k + 1, dtype=torch.float64, device=scatter.device)
    ws_bytes = lib.irp_pca_fit_workspace_bytes(d, k)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=scatter.device)
    info = (C.c_int32 * 4)()
    _lib.check(lib.irp_pca_fit_ex(_ptr(count), _ptr(total), _ptr(scatter), _ptr(shift), d, k, _ptr(mean), _ptr(comps),
                                  _ptr(evals), _ptr(ws), ws_bytes, int(solver), int(lanczos_first_check), info,
                                  _stream(count)), "irp_pca_fit_ex")
    return mean, comps, evals, {"solver": info[0], "lanczos_steps": info[1], "checks": info[2]}


@pca_fit.register_fake
def _(count, total, scatter, shift, k):
    d = scatter.shape[0]
    return (scatter.new_empty(d), scatter.new_empty((k, d)), scatter.new_empty(k + 1))


@torch.library.custom_op("irp_b200::pca_transform", mutates_args=())
@_on_device_of(0)
def pca_transform(x: torch.Tensor, mean: torch.Tensor, components: torch.Tensor) -> torch.Tensor:
    """(x - mean) @ components^T on the tensor cores (split-bf16 operands, fp32 accumulation) -> fp32 [n,k]."""
    lib = _lib_for(x)
    assert x.dtype == torch.float32 and x.is_contiguous()
    n, d = x.shape
    k = components.shape[0]
    z = torch.empty((n, k), dtype=torch.float32, device=x.device)
    ws_bytes = lib.irp_pca_transform_workspace_bytes(n, d, k)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=x.device)
    _lib.check(lib.irp_pca_transform(_ptr(x), n, d, _ptr(mean), _ptr(components.contiguous()), k, _ptr(z), _ptr(ws),
                                     ws_bytes, _stream(x)), "irp_pca_transform")
    return z


@pca_transform.register_fake
def _(x, mean, components):
    return x.new_empty((x.shape[0], components.shape[0]), dtype=torch.float32)


# ---------------------------------------------------------------------------------------------------------------
# A4 scoring
# ---------------------------------------------------------------------------------------------------------------
@torch.library.custom_op("irp_b200::lof", mutates_args=())
@_on_device_of(0)
def lof(z: torch.Tensor, group: Optional[torch.Tensor], n_groups: int, n_neighbors: int,
        contamination: float) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """-> (negative_outlier_factor fp64[n], offset fp64[n_groups], flags uint8[n])."""
    lib = _lib_for(z)
    assert z.dtype == torch.float32 and z.is_contiguous()
    n, d = z.shape
    scores = torch.empty(n, dtype=torch.float64, device=z.device)
    offsets = torch.empty(n_groups, dtype=torch.float64, device=z.device)
    flags = torch.empty(n, dtype=torch.uint8, device=z.device)
    ws_bytes = lib.irp_lof_workspace_bytes(n, d, n_neighbors)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=z.device)
    _lib.check(lib.irp_lof(_ptr(z), n, d, _ptr(group), n_groups, n_neighbors, C.c_double(contamination), _ptr(scores),
                           _ptr(offsets), _ptr(flags), _ptr(ws), ws_bytes, _stream(z)), "irp_lof")
    return scores, offsets, flags


@lof.register_fake
def _(z, group, n_groups, n_neighbors, contamination):
    n = z.shape[0]
    return (z.new_empty(n, dtype=torch.float64), z.new_empty(n_groups, dtype=torch.float64),
            z.new_empty(n, dtype=torch.uint8))


_side_streams = {}


def _side_stream(device: torch.device) -> torch.cuda.Stream:
    """One extra stream per device for work that runs beside the caller's stream (created on first use)."""
    key = device.index if device.index is not None else torch.cuda.current_device()
    if key not in _side_streams:
        _side_streams[key] = torch.cuda.Stream(device=device)
    return _side_streams[key]


@_on_device_of(0)
def lof_sharded_multi(z: torch.Tensor, problems, part: int, n_parts: int, all_reduce):
    """Several irp_lof problems over the SAME rows (e.g. per-class and global scoring) with the O(n^2) neighbour
    search sharded over `n_parts` ranks (irp_lof_knn_part / _lrd_part / _score_part / _finish).

    `problems` is a list of (group or None, n_groups, n_neighbors, contamination).  Every rank passes the same `z`
    / groups (all rows); `all_reduce(t)` must sum the fp64 tensor `t` in place over the ranks.  The problems advance
    in lockstep, so each of the three exchange steps is ONE all-reduce of a [len(problems), n] buffer.  Returns a
    list of (scores, offsets, flags), identical on every rank."""
    lib = _lib_for(z)
    assert z.dtype == torch.float32 and z.is_contiguous()
    n, d = z.shape
    np_ = len(problems)
    ws = []
    for (_, _, k, _) in problems:
        nbytes = lib.irp_lof_workspace_bytes(n, d, int(k))
        ws.append((torch.empty(nbytes, dtype=torch.uint8, device=z.device), nbytes))
    buf = lambda: torch.empty((np_, n), dtype=torch.float64, device=z.device)
    kdist, lrd, score = buf(), buf(), buf()
    st = _stream(z)
    # The neighbour searches (the O(n^2) phase) of the problems are independent: all but the last run on a side stream
    # beside the last one, and the caller's stream joins them before the first exchange.
    main = torch.cuda.current_stream(z.device)
    side = _side_stream(z.device) if np_ > 1 else None
    if side is not None:
        fork = torch.cuda.Event()
        fork.record(main)  # after every allocation above
        side.wait_event(fork)
    for i, (group, n_groups, k, _) in enumerate(problems):
        s_i = side if (side is not None and i < np_ - 1) else main
        with torch.cuda.stream(s_i):
            _lib.check(lib.irp_lof_knn_part(_ptr(z), n, d, _ptr(group), n_groups, int(k), part, n_parts,
                                            _ptr(kdist[i]), _ptr(ws[i][0]), ws[i][1], C.c_void_p(s_i.cuda_stream)),
                       "irp_lof_knn_part")
    if side is not None:
        joined = torch.cuda.Event()
        joined.record(side)
        main.wait_event(joined)
    all_reduce(kdist)
    for i, (_, n_groups, k, _) in enumerate(problems):
        _lib.check(lib.irp_lof_lrd_part(n, n_groups, int(k), part, n_parts, _ptr(kdist[i]), _ptr(lrd[i]),
                                        _ptr(ws[i][0]), ws[i][1], st), "irp_lof_lrd_part")
    all_reduce(lrd)
    for i, (_, n_groups, k, _) in enumerate(problems):
        _lib.check(lib.irp_lof_score_part(n, n_groups, int(k), part, n_parts, _ptr(lrd[i]), _ptr(score[i]),
                                          _ptr(ws[i][0]), ws[i][1], st), "irp_lof_score_part")
    all_reduce(score)
    out = []
    for i, (_, n_groups, k, contamination) in enumerate(problems):
        scores = torch.empty(n, dtype=torch.float64, device=z.device)
        offsets = torch.empty(n_groups, dtype=torch.float64, device=z.device)
        flags = torch.empty(n, dtype=torch.uint8, device=z.device)
        _lib.check(lib.irp_lof_finish(n, n_groups, int(k), C.c_double(contamination), _ptr(score[i]), _ptr(scores),
                                      _ptr(offsets), _ptr(flags), _ptr(ws[i][0]), ws[i][1], st), "irp_lof_finish")
        out.append((scores, offsets, flags))
    return out


def lof_sharded(z: torch.Tensor, group: Optional[torch.Tensor], n_groups: int, n_neighbors: int,
                contamination: float, part: int, n_parts: int, all_reduce) -> Tuple[torch.Tensor, torch.Tensor,
                                                                                   torch.Tensor]:
    """One irp_lof problem with the neighbour search sharded over `n_parts` ranks (see lof_sharded_multi)."""
    return lof_sharded_multi(z, [(group, n_groups, n_neighbors, contamination)], part, n_parts, all_reduce)[0]


@torch.library.custom_op("irp_b200::centroid_zscore", mutates_args=())
@_on_device_of(0)
def centroid_zscore(z: torch.Tensor, group: Optional[torch.Tensor], n_groups: int,
                    contamination: float) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor]:
    """-> (distance fp64[n], zscore fp64[n], threshold fp64[n_groups], flags uint8[n])."""
    lib = _lib_for(z)
    assert z.dtype == torch.float32 and z.is_contiguous()
    n, d = z.shape
    dist = torch.empty(n, dtype=torch.float64, device=z.device)
    zs = torch.empty(n, dtype=torch.float64, device=z.device)
    thr = torch.empty(n_groups, dtype=torch.float64, device=z.device)
    flags = torch.empty(n, dtype=torch.uint8, device=z.device)
    ws_bytes = lib.irp_centroid_workspace_bytes(n, d, n_groups)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=z.device)
    _lib.check(lib.irp_centroid_zscore(_ptr(z), n, d, _ptr(group), n_groups, C.c_double(contamination), _ptr(dist),
                                       _ptr(zs), _ptr(thr), _ptr(flags), _ptr(ws), ws_bytes, _stream(z)),
               "irp_centroid_zscore")
    return dist, zs, thr, flags


@centroid_zscore.register_fake
def _(z, group, n_groups, contamination):
    n = z.shape[0]
    return (z.new_empty(n, dtype=torch.float64), z.new_empty(n, dtype=torch.float64),
            z.new_empty(n_groups, dtype=torch.float64), z.new_empty(n, dtype=torch.uint8))


# ---------------------------------------------------------------------------------------------------------------
# N1 classifier head (functions/model.py:29-40) and the evaluate_full statistics (functions/train.py:208-216)
# ---------------------------------------------------------------------------------------------------------------
@torch.library.custom_op("irp_b200::classifier_head", mutates_args=())
@_on_device_of(0)
def classifier_head(features: torch.Tensor, w1: torch.Tensor, b1: torch.Tensor, w2: torch.Tensor,
                    b2: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """fp32 [B,in] features -> (logits fp32 [B,C], argmax int32 [B]) through Linear-ReLU-Linear (eval mode)."""
    lib = _lib_for(features)
    for t in (features, w1, b1, w2, b2):
        assert t.dtype == torch.float32 and t.is_contiguous() and t.is_cuda
    b, in_dim = features.shape
    hidden, c = w1.shape[0], w2.shape[0]
    assert w1.shape[1] == in_dim and w2.shape[1] == hidden and b1.numel() == hidden and b2.numel() == c
    logits = torch.empty((b, c), dtype=torch.float32, device=features.device)
    pred = torch.empty((b,), dtype=torch.int32, device=features.device)
    ws_bytes = lib.irp_classifier_head_workspace_bytes(b, hidden)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=features.device)
    _lib.check(lib.irp_classifier_head(_ptr(features), b, in_dim, _ptr(w1), _ptr(b1), hidden, _ptr(w2), _ptr(b2), c,
                                       _ptr(logits), _ptr(pred), _ptr(ws), ws_bytes, _stream(features)),
               "irp_classifier_head")
    return logits, pred


@classifier_head.register_fake
def _(features, w1, b1, w2, b2):
    b, c = features.shape[0], w2.shape[0]
    return features.new_empty((b, c)), features.new_empty((b,), dtype=torch.int32)


@_on_device_of(0)
def cross_entropy_stats(logits: torch.Tensor, labels: torch.Tensor,
                        class_weights: Optional[torch.Tensor] = None) -> torch.Tensor:
    """fp64 [3] on the device: sum of weighted per-row cross-entropies, sum of weights, number of correct rows."""
    lib = _lib_for(logits)
    assert logits.dtype == torch.float32 and logits.is_contiguous() and labels.dtype == torch.int64
    stats = torch.empty(3, dtype=torch.float64, device=logits.device)
    w = None
    if class_weights is not None:
        w = class_weights.to(device=logits.device, dtype=torch.float32).contiguous()
    _lib.check(lib.irp_cross_entropy_stats(_ptr(logits), _ptr(labels.contiguous()), logits.shape[0], logits.shape[1],
                                           _ptr(w), _ptr(stats), _stream(logits)), "irp_cross_entropy_stats")
    return stats


# ---------------------------------------------------------------------------------------------------------------
# N4 UMAP graph construction: exact k-NN arrays + fuzzy simplicial set weights
# ---------------------------------------------------------------------------------------------------------------
@torch.library.custom_op("irp_b200::knn_graph", mutates_args=())
@_on_device_of(0)
def knn_graph(z: torch.Tensor, n_neighbors: int) -> Tuple[torch.Tensor, torch.Tensor]:
    """umap_.py nearest_neighbors, exact: -> (knn_indices int32 [n,k], knn_dists float32 [n,k]); column 0 is the
    row itself at distance 0, then its k-1 nearest other rows in ascending (distance, index) order."""
    lib = _lib_for(z)
    assert z.dtype == torch.float32 and z.is_contiguous() and z.dim() == 2
    n, d = z.shape
    idx = torch.empty((n, n_neighbors), dtype=torch.int32, device=z.device)
    dist = torch.empty((n, n_neighbors), dtype=torch.float32, device=z.device)
    ws_bytes = lib.irp_knn_graph_workspace_bytes(n, d, n_neighbors)
    ws = torch.empty(max(ws_bytes, 8), dtype=torch.uint8, device=z.device)
    _lib.check(lib.irp_knn_graph(_ptr(z), n, d, n_neighbors, _ptr(idx), _ptr(dist), _ptr(ws), ws_bytes, _stream(z)),
               "irp_knn_graph")
    return idx, dist


@knn_graph.register_fake
def _(z, n_neighbors):
    n = z.shape[0]
    return z.new_empty((n, n_neighbors), dtype=torch.int32), z.new_empty((n, n_neighbors), dtype=torch.float32)


@torch.library.custom_op("irp_b200::umap_fuzzy_weights", mutates_args=())
@_on_device_of(0)
def umap_fuzzy_weights(knn_indices: torch.Tensor, knn_dists: torch.Tensor, local_connectivity: float = 1.0,
                       bandwidth: float = 1.0, n_iter: int = 64) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """umap_.py smooth_knn_dist + compute_membership_strengths: -> (sigmas f32 [n], rhos f32 [n], vals f32 [n,k]);
    vals[i, j] is the weight of the directed edge i -> knn_indices[i, j]."""
    lib = _lib_for(knn_dists)
    assert knn_indices.dtype == torch.int32 and knn_dists.dtype == torch.float32
    assert knn_indices.is_contiguous() and knn_dists.is_contiguous() and knn_indices.shape == knn_dists.shape
    n, k = knn_dists.shape
    sig = torch.empty(n, dtype=torch.float32, device=knn_dists.device)
    rho = torch.empty(n, dtype=torch.float32, device=knn_dists.device)
    vals = torch.empty((n, k), dtype=torch.float32, device=knn_dists.device)
    ws = torch.empty(8, dtype=torch.uint8, device=knn_dists.device)
    _lib.check(lib.irp_umap_fuzzy_weights(_ptr(knn_indices), _ptr(knn_dists), n, k, C.c_float(local_connectivity),
                                          C.c_float(bandwidth), int(n_iter), _ptr(sig), _ptr(rho), _ptr(vals),
                                          _ptr(ws), 8, _stream(knn_dists)), "irp_umap_fuzzy_weights")
    return sig, rho, vals


@umap_fuzzy_weights.register_fake
def _(knn_indices, knn_dists, local_connectivity=1.0, bandwidth=1.0, n_iter=64):
    n, k = knn_dists.shape
    return (knn_dists.new_empty(n), knn_dists.new_empty(n), knn_dists.new_empty((n, k)))
