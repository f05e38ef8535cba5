"""ORACLE (test infrastructure, never on the product path): numpy restatement of the graph-construction half of the
supervised UMAP the reference runs at /root/reference/functions/data_curation.py:704-705
(``umap.UMAP(**umap_params).fit_transform(features_pca, y=y_numeric)``), SURVEY.md section 8f row N4.

PARITY UNPINNED.  The arithmetic lives in umap-learn 0.5.7 (requirements.txt:175; pynndescent 0.5.13, numba 0.61.2),
which is neither vendored under /root/reference nor installable in the build container, and the reference holds no
test or golden vector for it.  The functions below restate that package's published algorithm (``umap/umap_.py``):

``nearest_neighbors``            k-NN arrays, the sample itself in column 0.  umap-learn searches EXACTLY
                                 (pairwise_distances + argsort, ``fast_knn_indices``) below 4 096 samples and with
                                 NN-descent above; this restatement (and the device path) is always exact
``smooth_knn_dist``              rho_i = distance to the local_connectivity-th nearest neighbour (interpolated),
                                 sigma_i by 64 steps of bisection so that sum_{j>=1} exp(-max(d_ij - rho_i, 0) /
                                 sigma_i) = log2(k) * bandwidth, floored at 1e-3 * the mean distance; float32 like
                                 the numba kernel's declared locals
``compute_membership_strengths`` directed weights exp(-(d_ij - rho_i) / sigma_i), 1 inside rho, 0 for the sample itself
``fuzzy_simplicial_set``         coo_matrix, A + A^T - A o A^T (set_op_mix_ratio interpolates with A o A^T)

What pins it here: tests/test_oracle.py checks the neighbour arrays against scikit-learn's brute-force
NearestNeighbors run in-process, and the weights against the equations above (the bisection's fixed point, rho, value
ranges, symmetry).  Anchors in the reference: the call site :704-705 and the default ``umap_params`` :688-694
(n_neighbors is UMAP's default 15, metric euclidean).
"""
from __future__ import annotations

import numpy as np

from . import lof_ref

SMOOTH_K_TOLERANCE = 1e-5
MIN_K_DIST_SCALE = 1e-3


def nearest_neighbors(x: np.ndarray, n_neighbors: int):
    """(knn_indices int32 [n,k], knn_dists float32 [n,k]): the row itself, then its k-1 nearest other rows."""
    x = np.asarray(x)
    n = x.shape[0]
    dist, idx = lof_ref.knn_bruteforce(x, n_neighbors - 1)
    knn_i = np.concatenate([np.arange(n, dtype=np.int64)[:, None], idx], 1).astype(np.int32)
    knn_d = np.concatenate([np.zeros((n, 1)), dist], 1).astype(np.float32)
    return knn_i, knn_d


def smooth_knn_dist(distances: np.ndarray, k: float, n_iter: int = 64, local_connectivity: float = 1.0,
                    bandwidth: float = 1.0):
    """(sigmas, rhos), both float32 [n].  Row loop of umap_.py smooth_knn_dist, vectorised over the rows."""
    f32 = np.float32
    d = np.asarray(distances, f32)
    n, kk = d.shape
    target = f32(np.log2(k) * bandwidth)
    rho = np.zeros(n, f32)
    mean_distances = f32(d.mean(dtype=np.float64))
    index = int(np.floor(local_connectivity))
    interpolation = f32(local_connectivity - index)
    for i in range(n):
        nz = d[i][d[i] > 0.0]
        if nz.shape[0] >= local_connectivity:
            if index > 0:
                rho[i] = nz[index - 1]
                if interpolation > SMOOTH_K_TOLERANCE:
                    rho[i] += interpolation * (nz[index] - nz[index - 1])
            else:
                rho[i] = interpolation * nz[0]
        elif nz.shape[0] > 0:
            rho[i] = nz.max()
    lo = np.zeros(n, f32)
    hi = np.full(n, np.inf, f32)
    mid = np.ones(n, f32)
    done = np.zeros(n, bool)
    dd = d[:, 1:] - rho[:, None]
    for _ in range(n_iter):
        with np.errstate(over="ignore", divide="ignore", invalid="ignore"):
            e = np.where(dd > 0, np.exp(-(dd / mid[:, None]), dtype=f32), f32(1.0))
        psum = e.sum(1, dtype=f32)
        done |= np.abs(psum - target) < SMOOTH_K_TOLERANCE
        if done.all():
            break
        gt = (psum > target) & ~done
        le = ~gt & ~done
        hi = np.where(gt, mid, hi)
        new_mid_gt = (lo + hi) / f32(2.0)
        lo = np.where(le, mid, lo)
        new_mid_le = np.where(np.isinf(hi), mid * f32(2.0), (lo + hi) / f32(2.0))
        mid = np.where(gt, new_mid_gt, np.where(le, new_mid_le, mid)).astype(f32)
    result = mid.copy()
    mean_row = d.mean(1, dtype=np.float64).astype(f32)
    floor = np.where(rho > 0.0, MIN_K_DIST_SCALE * mean_row, MIN_K_DIST_SCALE * mean_distances).astype(f32)
    result = np.maximum(result, floor)
    return result.astype(f32), rho


def compute_membership_strengths(knn_indices, knn_dists, sigmas, rhos):
    """(rows, cols, vals) of the directed graph, float32 vals."""
    f32 = np.float32
    idx = np.asarray(knn_indices)
    d = np.asarray(knn_dists, f32)
    n, k = idx.shape
    rows = np.repeat(np.arange(n, dtype=np.int32), k)
    cols = idx.reshape(-1).astype(np.int32)
    diff = d - np.asarray(rhos, f32)[:, None]
    sig = np.asarray(sigmas, f32)[:, None]
    with np.errstate(divide="ignore", invalid="ignore", over="ignore"):
        vals = np.where((diff <= 0.0) | (sig == 0.0), f32(1.0), np.exp(-(diff / sig), dtype=f32))
    vals = np.where(idx == np.arange(n)[:, None], f32(0.0), vals)
    vals = np.where(idx < 0, f32(0.0), vals)
    return rows, cols, vals.reshape(-1).astype(f32)


def fuzzy_simplicial_set(x, n_neighbors: int, set_op_mix_ratio: float = 1.0, local_connectivity: float = 1.0,
                         knn_indices=None, knn_dists=None):
    """(graph coo/csr [n,n], sigmas, rhos) like umap_.py fuzzy_simplicial_set(metric="euclidean")."""
    import scipy.sparse

    if knn_indices is None or knn_dists is None:
        knn_indices, knn_dists = nearest_neighbors(x, n_neighbors)
    sigmas, rhos = smooth_knn_dist(knn_dists, float(n_neighbors), local_connectivity=float(local_connectivity))
    rows, cols, vals = compute_membership_strengths(knn_indices, knn_dists, sigmas, rhos)
    n = knn_indices.shape[0]
    keep = cols >= 0
    result = scipy.sparse.coo_matrix((vals[keep], (rows[keep], cols[keep])), shape=(n, n))
    result.eliminate_zeros()
    transpose = result.transpose()
    prod_matrix = result.multiply(transpose)
    result = set_op_mix_ratio * (result + transpose - prod_matrix) + (1.0 - set_op_mix_ratio) * prod_matrix
    result.eliminate_zeros()
    return result, sigmas, rhos
