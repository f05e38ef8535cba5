"""Developer tool: UMAP graph construction on PCA-shaped rows -- irp_knn_graph +
irp_umap_fuzzy_weights on the device against scikit-learn's brute-force neighbours (all host cores) + the numpy
restatement of umap-learn's smooth_knn_dist on the host (umap-learn itself is not installable here)."""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "image-recognition-pipeline_b200"))
sys.path.insert(0, ROOT)
from irp_b200 import ops  # noqa: E402
from oracle import umap_graph_ref as ug  # noqa: E402

k = 15
for n, d in [(27000, 50), (125000, 128)]:
    rng = np.random.default_rng(0)
    centers = rng.normal(size=(10, d)).astype(np.float32) * 4
    x = (centers[rng.integers(0, 10, n)] + rng.normal(size=(n, d)).astype(np.float32)).astype(np.float32)
    xt = torch.from_numpy(x).cuda()
    for _ in range(2):
        idx, dist = ops.knn_graph(xt, k)
        sig, rho, vals = ops.umap_fuzzy_weights(idx, dist)
    torch.cuda.synchronize()
    e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    e[0].record()
    idx, dist = ops.knn_graph(xt, k)
    e[1].record()
    sig, rho, vals = ops.umap_fuzzy_weights(idx, dist)
    e[2].record()
    torch.cuda.synchronize()
    line = {"rows": n, "dim": d, "n_neighbors": k, "knn_graph_ms": e[0].elapsed_time(e[1]),
            "fuzzy_weights_ms": e[1].elapsed_time(e[2])}
    if n <= 30000:
        from sklearn.neighbors import NearestNeighbors
        t0 = time.perf_counter()
        dd, ii = NearestNeighbors(n_neighbors=k, algorithm="brute", n_jobs=-1).fit(x).kneighbors(x)
        t1 = time.perf_counter()
        dd[:, 0] = 0.0  # sklearn's expanded-form float32 distances leave ~1e-3 on the diagonal
        sr, rr = ug.smooth_knn_dist(dd.astype(np.float32), float(k))
        t2 = time.perf_counter()
        line.update({"cpu_sklearn_brute_knn_ms": (t1 - t0) * 1e3, "cpu_numpy_smooth_knn_ms": (t2 - t1) * 1e3,
                     "cpu_cores": os.cpu_count(),
                     "max_dist_diff": float(np.abs(dist.cpu().numpy() - dd).max()),
                     "sigma_rel_diff": float(np.abs(sig.cpu().numpy() - sr).max() / sr.max())})
    print(json.dumps(line), flush=True)
