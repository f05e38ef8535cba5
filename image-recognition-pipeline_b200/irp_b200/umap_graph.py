"""Host-side mirror of the graph-construction functions of umap-learn 0.5.7 (``umap/umap_.py``) that the supervised
UMAP of ``create_embeddings`` (/root/reference/functions/data_curation.py:704-705) runs first, on the library's
kernels:

``nearest_neighbors``      -> ``irp_knn_graph``: exact Euclidean k-NN arrays (the row itself first), on the device
``fuzzy_simplicial_set``   -> ``irp_knn_graph`` + ``irp_umap_fuzzy_weights`` on the device, then the sparse
                              symmetrisation ``A + A^T - A o A^T`` with scipy on the host (same return values as
                              umap's function: graph, sigmas, rhos)

UMAP itself stays host-side (north_star): ``functions.data_curation.create_embeddings`` hands the arrays to
``umap.UMAP(precomputed_knn=(knn_indices, knn_dists))`` so that UMAP skips its own neighbour search.  There is no CPU
implementation behind these functions.  parity unpinned: umap-learn is not installable in the build container.
"""
from __future__ import annotations

import numpy as np
import torch

from . import ops


def _device(device=None) -> torch.device:
    if device is None:
        return torch.device(f"cuda:{torch.cuda.current_device()}")
    return torch.device(device)


def nearest_neighbors(X, n_neighbors: int, device=None):
    """(knn_indices int32 [n,k], knn_dists float32 [n,k]) of the rows of X under the Euclidean metric, exact."""
    z = torch.from_numpy(np.ascontiguousarray(X, dtype=np.float32)).to(_device(device))
    idx, dist = ops.knn_graph(z, int(n_neighbors))
    return idx.cpu().numpy(), dist.cpu().numpy()


def fuzzy_simplicial_set(X, n_neighbors: int, knn_indices=None, knn_dists=None, set_op_mix_ratio: float = 1.0,
                         local_connectivity: float = 1.0, apply_set_operations: bool = True, device=None):
    """umap_.py fuzzy_simplicial_set for the Euclidean metric -> (graph scipy.sparse.coo_matrix [n,n], sigmas, rhos)."""
    import scipy.sparse

    dev = _device(device)
    if knn_indices is None or knn_dists is None:
        z = torch.from_numpy(np.ascontiguousarray(X, dtype=np.float32)).to(dev)
        idx_t, dist_t = ops.knn_graph(z, int(n_neighbors))
    else:
        idx_t = torch.from_numpy(np.ascontiguousarray(knn_indices, dtype=np.int32)).to(dev)
        dist_t = torch.from_numpy(np.ascontiguousarray(knn_dists, dtype=np.float32)).to(dev)
    sig_t, rho_t, val_t = ops.umap_fuzzy_weights(idx_t, dist_t, float(local_connectivity), 1.0, 64)
    idx, vals = idx_t.cpu().numpy(), val_t.cpu().numpy()
    n, k = idx.shape
    rows = np.repeat(np.arange(n, dtype=np.int32), k)
    cols = idx.reshape(-1)
    keep = cols >= 0
    result = scipy.sparse.coo_matrix((vals.reshape(-1)[keep], (rows[keep], cols[keep])), shape=(n, n))
    result.eliminate_zeros()
    if apply_set_operations:
        transpose = result.transpose()
        prod_matrix = result.multiply(transpose)
        result = set_op_mix_ratio * (result + transpose - prod_matrix) + (1.0 - set_op_mix_ratio) * prod_matrix
    result.eliminate_zeros()
    return result, sig_t.cpu().numpy(), rho_t.cpu().numpy()
