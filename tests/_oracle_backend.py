"""Test-only backend for OutlierStage: every compute step is the numpy oracle, on CPU tensors.

Lets the CPU suite exercise the host-side logic of the stage (batching, sharding, the single all-reduce of the
PCA partial sums, the gather of projected rows) under gloo without a GPU.  Never imported by the product."""
import numpy as np
import torch

from oracle import lof_ref, pil_resample


class OracleBackend:
    device = torch.device("cpu")

    def __init__(self, embed_dim=64, seed=0):
        rng = np.random.default_rng(seed)
        # stand-in for the trunk: fixed random projection of a 16x16 average-pooled preprocessed image
        self.proj = rng.standard_normal((3 * 16 * 16, embed_dim)).astype(np.float32) / 10.0
        self.embed_calls = 0

    def embed(self, part, max_taps):
        self.embed_calls += 1
        feats = []
        pix = part.pixels.numpy()
        for (h, w), off in zip(part.hw_np, part.offsets_np):
            img = pix[off:off + h * w * 3].reshape(h, w, 3)
            x = pil_resample.transform(img)  # [3,224,224]
            pooled = x.reshape(3, 16, 14, 16, 14).mean(axis=(2, 4)).reshape(-1)
            feats.append(np.maximum(pooled @ self.proj + 1.0, 0))
        return torch.from_numpy(np.stack(feats).astype(np.float32))

    @staticmethod
    def cov_accumulate(x, shift, count, total, scatter):
        y = x.numpy().astype(np.float64) - shift.numpy().astype(np.float64)
        count += y.shape[0]
        total += torch.from_numpy(y.sum(0))
        scatter += torch.from_numpy(y.T @ y)

    @staticmethod
    def pca_fit(count, total, scatter, shift, k):
        n = float(count.item())
        delta = total.numpy() / n
        cov = (scatter.numpy() - n * np.outer(delta, delta)) / (n - 1)
        evals, evecs = np.linalg.eigh(cov)
        evals, evecs = evals[::-1], evecs[:, ::-1]
        comps = evecs[:, :k].T.copy()
        idx = np.argmax(np.abs(comps), axis=1)
        comps *= np.sign(comps[np.arange(k), idx])[:, None]
        out = np.concatenate([np.maximum(evals[:k], 0), [np.trace(cov)]])
        return (torch.from_numpy(shift.numpy().astype(np.float64) + delta), torch.from_numpy(comps),
                torch.from_numpy(out))

    @staticmethod
    def pca_transform(x, mean, comps):
        z = (x.numpy().astype(np.float64) - mean.numpy()) @ comps.numpy().T
        return torch.from_numpy(z.astype(np.float32))

    @staticmethod
    def lof(z, group, n_groups, k, contamination):
        zn = z.numpy()
        g = np.zeros(len(zn), np.int64) if group is None else group.numpy().astype(np.int64)
        scores = np.zeros(len(zn))
        offsets = np.zeros(n_groups)
        flags = np.zeros(len(zn), np.uint8)
        for c in range(n_groups):
            m = g == c
            if m.sum() < 2:
                continue
            f, s, o = lof_ref.lof_fit_predict(zn[m], k, contamination)
            scores[m], offsets[c], flags[m] = s, o, f
        return torch.from_numpy(scores), torch.from_numpy(offsets), torch.from_numpy(flags)

    @staticmethod
    def lof_sharded_multi(z, problems, part, n_parts, all_reduce):
        """Same protocol as ops.lof_sharded_multi (irp_lof_knn_part / _lrd_part / _score_part / _finish): this rank
        does the neighbour search for the rows it owns (row position % n_parts inside its group); the problems
        advance in lockstep and each of the three [problems, n] fp64 buffers is summed across ranks ONCE."""
        zn = z.numpy()
        n = len(zn)
        npb = len(problems)
        kdist, lrd, score = (torch.zeros((npb, n), dtype=torch.float64) for _ in range(3))
        groups, nbrs = [], []
        for pi, (group, n_groups, k, _) in enumerate(problems):
            g = np.zeros(n, np.int64) if group is None else group.numpy().astype(np.int64)
            groups.append(g)
            nbr = {}
            for c in range(n_groups):
                rows = np.flatnonzero(g == c)
                if len(rows) < 2:
                    continue
                kk = max(1, min(k, len(rows) - 1))
                sel = np.arange(len(rows)) % n_parts == part
                mine = rows[sel]
                dist, idx = lof_ref.knn_bruteforce(zn[rows], kk)
                if zn.dtype == np.float32:
                    dist = dist.astype(np.float32).astype(np.float64)
                nbr[c] = (rows, kk, dist[sel], rows[idx[sel]], mine)
                kdist[pi, mine] = torch.from_numpy(dist[sel, kk - 1])
            nbrs.append(nbr)
        all_reduce(kdist)
        for pi, nbr in enumerate(nbrs):
            kd = kdist[pi].numpy()
            for c, (rows, kk, dist, idx, mine) in nbr.items():
                reach = np.maximum(dist, kd[idx])
                lrd[pi, mine] = torch.from_numpy(1.0 / (reach.mean(axis=1) + 1e-10))
        all_reduce(lrd)
        for pi, nbr in enumerate(nbrs):
            lr = lrd[pi].numpy()
            for c, (rows, kk, dist, idx, mine) in nbr.items():
                score[pi, mine] = torch.from_numpy(-(lr[idx] / lr[mine][:, None]).mean(axis=1))
        all_reduce(score)
        out = []
        for pi, (group, n_groups, k, contamination) in enumerate(problems):
            sc = score[pi].numpy()
            g = groups[pi]
            offsets = np.zeros(n_groups)
            flags = np.zeros(n, np.uint8)
            for c in range(n_groups):
                m = g == c
                if m.sum() < 2:
                    continue
                offsets[c] = np.percentile(sc[m], 100.0 * contamination)
                flags[m] = sc[m] < offsets[c]
            out.append((torch.from_numpy(sc.copy()), torch.from_numpy(offsets), torch.from_numpy(flags)))
        return out

    @classmethod
    def lof_sharded(cls, z, group, n_groups, k, contamination, part, n_parts, all_reduce):
        return cls.lof_sharded_multi(z, [(group, n_groups, k, contamination)], part, n_parts, all_reduce)[0]
