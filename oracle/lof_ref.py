"""ORACLE (test infrastructure, never on the product path): numpy restatement of the reference's outlier scoring.

``detect_outliers`` (/root/reference/functions/data_curation.py:709-728) label-encodes the classes, runs
``LocalOutlierFactor(n_neighbors=30, contamination=0.05).fit_predict`` on the rows of every class and
``LocalOutlierFactor(75, 0.03)`` on all rows, and returns ``== -1`` masks in input order.

LOF arithmetic restated from scikit-learn (pinned 1.6.1, container 1.9.0; not vendored under /root/reference):
sklearn/neighbors/_lof.py:286-293 (k clipped to n-1), :295-304 (k nearest neighbours excluding the sample itself;
distances cast to float32 when the input is float32), :498-523 (reachability = max(d, k-distance of the
neighbour), lrd = 1 / (mean reach + 1e-10)), :306-323 (score = -mean(lrd[nbrs] / lrd[i]), offset =
np.percentile(score, 100 * contamination), outlier iff score < offset).

The centroid / z-score scorer named by the north_star has NO counterpart in the reference; its
definition is written down here and this file is its only oracle ("parity unpinned" for that scorer).

Pinned: tests/test_oracle.py compares lof_scores / detect_outliers with sklearn run in-process and with
tests/golden/lof.npz, which was produced by the unmodified reference function (oracle/make_golden.py).
"""
from __future__ import annotations

import numpy as np


def label_encode(labels) -> tuple[np.ndarray, np.ndarray]:
    """sklearn LabelEncoder: classes = sorted unique labels, ids = position in classes (data_curation.py:712-713)."""
    classes, ids = np.unique(np.asarray(labels), return_inverse=True)
    return classes, ids.astype(np.int64)


def knn_bruteforce(z: np.ndarray, k: int, chunk: int = 2048) -> tuple[np.ndarray, np.ndarray]:
    """k nearest neighbours of every row among the OTHER rows (Euclidean, float64), sorted by distance."""
    z64 = np.asarray(z, np.float64)
    n = z64.shape[0]
    sq = np.einsum("ij,ij->i", z64, z64)
    dist = np.empty((n, k), np.float64)
    idx = np.empty((n, k), np.int64)
    for s in range(0, n, chunk):
        e = min(n, s + chunk)
        diff2 = sq[s:e, None] - 2.0 * (z64[s:e] @ z64.T) + sq[None, :]
        # exact recomputation is cheap at oracle sizes and avoids the cancellation of the expanded form
        if n * (e - s) * z64.shape[1] <= 2e8:
            diff2 = ((z64[s:e, None, :] - z64[None, :, :]) ** 2).sum(-1)
        np.maximum(diff2, 0.0, out=diff2)
        diff2[np.arange(e - s), np.arange(s, e)] = np.inf  # exclude the sample itself
        part = np.argpartition(diff2, k - 1, axis=1)[:, :k]
        pd = np.take_along_axis(diff2, part, axis=1)
        order = np.argsort(pd, axis=1, kind="stable")
        idx[s:e] = np.take_along_axis(part, order, axis=1)
        dist[s:e] = np.sqrt(np.take_along_axis(pd, order, axis=1))
    return dist, idx


def lof_scores(z: np.ndarray, n_neighbors: int) -> np.ndarray:
    """negative_outlier_factor_ of LocalOutlierFactor(n_neighbors).fit(z) (sklearn/neighbors/_lof.py:286-323)."""
    z = np.asarray(z)
    n = z.shape[0]
    k = max(1, min(n_neighbors, n - 1))
    dist, idx = knn_bruteforce(z, k)
    if z.dtype == np.float32:  # _lof.py:299-303
        dist = dist.astype(np.float32)
    dist_k = dist[idx, k - 1]
    reach = np.maximum(dist, dist_k)
    lrd = 1.0 / (np.mean(reach, axis=1) + 1e-10)
    ratios = lrd[idx] / lrd[:, None]
    return -np.mean(ratios, axis=1)


def lof_fit_predict(z: np.ndarray, n_neighbors: int, contamination: float):
    """(is_outlier bool[n], scores, offset) -- fit_predict(...) == -1 with a numeric contamination."""
    scores = lof_scores(z, n_neighbors)
    offset = np.percentile(scores, 100.0 * contamination)
    return scores < offset, scores, offset


def detect_outliers(embedding, labels, class_n_neighbors=30, class_contamination=0.05, global_n_neighbors=75,
                    global_contamination=0.03):
    """Restatement of data_curation.py:709-728: (class_outliers bool[n], global_outliers bool[n])."""
    embedding = np.asarray(embedding)
    _, y = label_encode(labels)
    class_out = np.zeros(len(y), dtype=bool)
    for c in np.unique(y):
        mask = y == c
        class_out[mask] = lof_fit_predict(embedding[mask], class_n_neighbors, class_contamination)[0]
    global_out = lof_fit_predict(embedding, global_n_neighbors, global_contamination)[0]
    return class_out, global_out


def centroid_zscore(z: np.ndarray, groups: np.ndarray, n_groups: int, contamination: float):
    """Distance-to-centroid scorer (no reference counterpart).

    Per group g: mu_g = mean of its rows; d_i = ||z_i - mu_g||_2; zscore_i = (d_i - mean_g d) / std_g d
    (population std; 0 when std is 0); threshold_g = np.percentile(d_g, 100 * (1 - contamination));
    flag_i = d_i > threshold_g.  Returns (dist, zscore, thresholds[n_groups], flags)."""
    z64 = np.asarray(z, np.float64)
    n = z64.shape[0]
    groups = np.zeros(n, np.int64) if groups is None else np.asarray(groups, np.int64)
    dist = np.zeros(n)
    zs = np.zeros(n)
    thr = np.full(n_groups, np.nan)
    flags = np.zeros(n, bool)
    for g in range(n_groups):
        m = groups == g
        if not m.any():
            continue
        mu = z64[m].mean(axis=0)
        d = np.sqrt(((z64[m] - mu) ** 2).sum(axis=1))
        sd = d.std()
        dist[m] = d
        zs[m] = (d - d.mean()) / sd if sd > 0 else 0.0
        thr[g] = np.percentile(d, 100.0 * (1.0 - contamination))
        flags[m] = d > thr[g]
    return dist, zs, thr, flags
