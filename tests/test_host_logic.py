"""CPU: host-side logic of the stage -- packing, slicing, sharding, batching, and the multi-rank path under gloo
(world_size 2) with the oracle-backed test backend."""
import os
import socket

import numpy as np
import pytest
import torch

from _oracle_backend import OracleBackend
from irp_b200 import stage as st
from oracle import pil_resample, synth


def _images(n, seed=0):
    rng = np.random.default_rng(seed)
    sizes = [(64, 80), (120, 90), (224, 224), (50, 50), (300, 140)]
    return [synth.smooth_image(rng, *sizes[i % len(sizes)], i % 3) for i in range(n)]


def test_pack_images_layout_and_slices():
    imgs = _images(7)
    p = st.pack_images(imgs, pin=False)
    assert len(p) == 7 and p.pixels.dtype == torch.uint8
    assert (p.offsets_np % st.ALIGN == 0).all()
    for i, im in enumerate(imgs):
        o = int(p.offsets_np[i])
        assert np.array_equal(p.pixels.numpy()[o:o + im.size].reshape(im.shape), im)
    s = p.slice(2, 5)
    assert len(s) == 3 and int(s.offsets_np[0]) == 0
    for j, im in enumerate(imgs[2:5]):
        o = int(s.offsets_np[j])
        assert np.array_equal(s.pixels.numpy()[o:o + im.size].reshape(im.shape), im)
    assert p.nbytes() == sum(im.size for im in imgs)


def test_pack_rejects_non_rgb():
    with pytest.raises(ValueError):
        st.pack_images([np.zeros((4, 4), np.uint8)], pin=False)
    with pytest.raises(ValueError):
        st.pack_images([np.zeros((4, 4, 3), np.float32)], pin=False)


def test_taps_bound_matches_oracle():
    rng = np.random.default_rng(1)
    for _ in range(500):
        h, w = int(rng.integers(30, 4000)), int(rng.integers(30, 4000))
        assert st.taps_for(h, w) == pil_resample.max_taps(h, w)


@pytest.mark.parametrize("n,ws", [(10, 3), (27000, 8), (5, 8), (0, 2), (1024, 1)])
def test_shard_range_partitions_exactly(n, ws):
    ranges = [st.shard_range(n, r, ws) for r in range(ws)]
    assert ranges[0][0] == 0 and ranges[-1][1] == n
    for (a, b), (c, d) in zip(ranges, ranges[1:]):
        assert b == c and a <= b
    sizes = [b - a for a, b in ranges]
    assert max(sizes) - min(sizes) <= 1


def _run_stage(images, ids, n_classes, k=6, pg=None, batch_size=4):
    stage = st.OutlierStage(OracleBackend(), batch_size=batch_size, pca_components=k, class_n_neighbors=5,
                            class_contamination=0.1, global_n_neighbors=8, global_contamination=0.1,
                            process_group=pg, embed_dim=64)
    packed = st.pack_images(images, pin=False)
    return stage, stage.run(packed, torch.from_numpy(ids), n_classes, from_host=True)


def test_single_rank_stage_batches_and_matches_direct_oracle():
    images = _images(22)
    ids = (np.arange(22) % 3).astype(np.int32)
    stage, res = _run_stage(images, ids, 3, batch_size=5)
    assert stage.backend.embed_calls == 5  # ceil(22 / 5)
    assert res.features.shape == (22, 64) and res.z.shape == (22, 6)
    # PCA state equals the oracle fit of the same features
    from oracle import pca_ref
    ref = pca_ref.pca_fit(res.features.numpy(), 6)
    assert pca_ref.subspace_angle(res.pca.components.numpy(), ref.components) < 1e-6
    np.testing.assert_allclose(res.pca.explained_variance.numpy(), ref.explained_variance, rtol=1e-6)
    np.testing.assert_allclose(res.pca.total_variance, ref.eigenvalues.sum(), rtol=1e-6)
    assert res.pca.n_samples == 22
    assert res.class_outliers.dtype == torch.bool and res.class_outliers.shape == (22,)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world_size, port, out_dir):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world_size)
    images = _images(23, seed=3)
    ids = (np.arange(23) % 3).astype(np.int32)
    lo, hi = st.shard_range(len(images), rank, world_size)
    _, res = _run_stage(images[lo:hi], ids[lo:hi], 3)
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), z=res.z.numpy(), comps=res.pca.components.numpy(),
             ev=res.pca.explained_variance.numpy(), cls=res.class_outliers.numpy(), glob=res.global_outliers.numpy(),
             n=res.pca.n_samples, feats=res.features.numpy())
    dist.destroy_process_group()


def test_two_rank_gloo_matches_single_rank(tmp_path):
    import torch.multiprocessing as mp
    mp.spawn(_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    images = _images(23, seed=3)
    ids = (np.arange(23) % 3).astype(np.int32)
    _, ref = _run_stage(images, ids, 3)
    r0, r1 = np.load(tmp_path / "rank0.npz"), np.load(tmp_path / "rank1.npz")
    assert int(r0["n"]) == int(r1["n"]) == 23
    assert r0["feats"].shape[0] + r1["feats"].shape[0] == 23
    from oracle import pca_ref
    for r in (r0, r1):
        # every rank holds the same global model and the same flags for all 23 rows (rank order = image order)
        assert pca_ref.subspace_angle(r["comps"], ref.pca.components.numpy()) < 1e-6
        np.testing.assert_allclose(r["ev"], ref.pca.explained_variance.numpy(), rtol=1e-7)
        np.testing.assert_allclose(r["z"], ref.z.numpy(), atol=1e-4)
        assert np.array_equal(r["cls"], ref.class_outliers.numpy())
        assert np.array_equal(r["glob"], ref.global_outliers.numpy())


def test_classifier_drop_ins_refuse_to_run_without_the_cuda_path():
    """No CPU fallback on the widened path either: evaluate_full wants a B200Classifier, and building one needs a
    CUDA device."""
    import torch
    from functions import train as b200_train
    from irp_b200.classifier import B200Classifier
    from oracle import classifier_ref
    model = classifier_ref.build_classifier(3, seed=0)
    with pytest.raises(TypeError):
        b200_train.evaluate_full(model, [], torch.nn.CrossEntropyLoss(), disable_progress=True)
    with pytest.raises(RuntimeError):
        B200Classifier(model, "cpu")
    with pytest.raises(ValueError):
        B200Classifier(torch.nn.Linear(4, 4), "cpu")


def test_bench_union_of_intervals():
    """bench.py times the trunk as the union of the lanes' call intervals (overlapping calls must not add up)."""
    import bench
    assert bench.union_ms([]) == 0.0
    assert bench.union_ms([(0.0, 1.0), (2.0, 3.5)]) == 2.5
    assert bench.union_ms([(0.0, 2.0), (1.0, 3.0), (2.5, 2.75)]) == 3.0      # overlapping / nested
    assert bench.union_ms([(5.0, 6.0), (0.0, 1.0), (0.5, 5.5)]) == 6.0       # unsorted input
