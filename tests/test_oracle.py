"""CPU: the oracle restatements against the golden vectors produced by the unmodified reference
(oracle/make_golden.py) and against the third-party libraries the reference calls, run in-process."""
import hashlib
import warnings

import numpy as np
import pytest

from conftest import load_golden
from oracle import lof_ref, pca_ref, pil_resample, stage_ref, synth


# ------------------------------------------------------------------ A1 preprocess
def _golden_image(seed, h, w):
    return np.random.default_rng(int(seed)).integers(0, 256, (int(h), int(w), 3), dtype=np.uint8)


def test_resample_matches_reference_golden_bit_for_bit():
    g = load_golden("preprocess.npz")
    for (h, w), seed, crop, sha in zip(g["sizes"], g["seeds"], g["crops"], g["float32_sha256"]):
        img = _golden_image(seed, h, w)
        u8 = pil_resample.transform_u8(img)
        assert np.array_equal(u8, crop), f"{h}x{w}: uint8 crop differs from the reference transform"
        f32 = pil_resample.normalize(u8)
        assert hashlib.sha256(f32.tobytes()).hexdigest() == str(sha), f"{h}x{w}: normalised float32 differs"


@pytest.mark.parametrize("h,w", [(480, 640), (1000, 700), (232, 232), (640, 232), (60, 57), (225, 1200)])
def test_resample_matches_pillow_and_torchvision_in_process(h, w):
    from PIL import Image
    from torchvision.models import ResNet50_Weights
    img = np.random.default_rng(h * 10007 + w).integers(0, 256, (h, w, 3), dtype=np.uint8)
    ref = ResNet50_Weights.DEFAULT.transforms()(Image.fromarray(img)).numpy()
    assert np.array_equal(pil_resample.transform(img), ref)


def test_geometry_matches_c_abi_host_helper():
    from irp_b200 import _lib
    rng = np.random.default_rng(0)
    for _ in range(300):
        h, w = int(rng.integers(40, 3000)), int(rng.integers(40, 3000))
        oh, ow = pil_resample.resized_size(h, w)
        top, left = pil_resample.crop_offsets(oh, ow)
        assert _lib.geometry(h, w) == (oh, ow, top, left, pil_resample.max_taps(h, w))


def test_bf16_rounding_helper_matches_torch():
    import torch
    x = np.random.default_rng(1).standard_normal(10000).astype(np.float32) * 3
    bits = pil_resample.to_bf16_bits(x)
    ref = torch.from_numpy(x).bfloat16().view(torch.int16).numpy().view(np.uint16)
    assert np.array_equal(bits, ref)


# ------------------------------------------------------------------ A2 embeddings
def test_stage_ref_embeddings_match_reference_golden():
    g = load_golden("embeddings.npz")
    hw = g["hw"]
    sizes = hw[:, 0].astype(np.int64) * hw[:, 1] * 3
    offs = np.concatenate([[0], np.cumsum(sizes)])
    images = [g["pixels"][offs[i]:offs[i + 1]].reshape(hw[i, 0], hw[i, 1], 3) for i in range(len(hw))]
    feats = stage_ref.embed_arrays(images, batch_size=1, seed=int(g["seed"]))
    ref = g["features"]
    # same library calls as the reference; batch-1 like the reference -> identical up to thread scheduling noise
    np.testing.assert_allclose(feats, ref, rtol=1e-4, atol=1e-4)
    cos = (feats * ref).sum(1) / np.linalg.norm(feats, axis=1) / np.linalg.norm(ref, axis=1)
    assert cos.min() > 0.999999


# ------------------------------------------------------------------ A3 PCA
def test_pca_ref_matches_golden_full_solver():
    g = load_golden("pca.npz")
    x = synth.embedding_like(int(g["n"]), int(g["d"]), seed=int(g["seed"]))
    k = int(g["k"])
    r = pca_ref.pca_fit(x, k)
    np.testing.assert_allclose(r.mean, g["mean"], rtol=0, atol=1e-10)
    np.testing.assert_allclose(r.explained_variance, g["explained_variance"], rtol=1e-9)
    np.testing.assert_allclose(r.explained_variance_ratio, g["explained_variance_ratio"], rtol=1e-9)
    np.testing.assert_allclose(r.singular_values, g["singular_values"], rtol=1e-9)
    np.testing.assert_allclose(r.noise_variance, float(g["noise_variance"]), rtol=1e-8)
    assert pca_ref.subspace_angle(r.components, g["components"]) < 1e-7
    # identical signs (svd_flip convention) and per-component agreement
    dots = (r.components * g["components"]).sum(1)
    assert (dots > 0.999999).all()
    z = pca_ref.pca_transform(x[:64], r.mean, r.components)
    np.testing.assert_allclose(z, g["z_head"], atol=1e-7)


def test_subspace_angle_helper():
    rng = np.random.default_rng(0)
    q = np.linalg.qr(rng.standard_normal((50, 5)))[0].T
    assert pca_ref.subspace_angle(q, q) < 1e-12
    rot = q.copy()
    theta = 1e-4
    rot[0] = np.cos(theta) * q[0] + np.sin(theta) * np.linalg.qr(rng.standard_normal((50, 6)))[0].T[5]
    assert 0 < pca_ref.subspace_angle(q, rot) < 2e-4


# ------------------------------------------------------------------ A4 LOF
def test_lof_ref_matches_reference_golden():
    g = load_golden("lof.npz")
    z, y = synth.clustered_points(int(g["n"]), int(g["d"]), int(g["classes"]), seed=int(g["seed"]))
    labels = np.array([f"cls{c:02d}" for c in y])
    scores = lof_ref.lof_scores(z, 75)
    np.testing.assert_allclose(scores, g["global_scores"], rtol=2e-6)
    cls_out, glob_out = lof_ref.detect_outliers(z, labels)
    assert np.array_equal(glob_out, g["global_outliers"])
    assert np.array_equal(cls_out, g["class_outliers"])


def test_lof_ref_matches_reference_golden_2d_clipped_k():
    g = load_golden("lof.npz")
    z, y = synth.clustered_points(int(g["n2"]), int(g["d2"]), int(g["classes2"]), seed=int(g["seed2"]))
    labels = np.array([f"c{c:02d}" for c in y])
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        cls_out, glob_out = lof_ref.detect_outliers(z, labels)
    assert np.array_equal(glob_out, g["global_outliers2"])
    assert np.array_equal(cls_out, g["class_outliers2"])


def test_stage_ref_detect_outliers_is_the_reference_call():
    g = load_golden("lof.npz")
    z, y = synth.clustered_points(int(g["n"]), int(g["d"]), int(g["classes"]), seed=int(g["seed"]))
    labels = np.array([f"cls{c:02d}" for c in y])
    cls_out, glob_out = stage_ref.detect_outliers(z, labels)
    assert np.array_equal(cls_out, g["class_outliers"]) and np.array_equal(glob_out, g["global_outliers"])


def test_centroid_scorer_definition():
    z, y = synth.clustered_points(500, 8, 4, seed=2)
    dist, zs, thr, flags = lof_ref.centroid_zscore(z, y, 4, 0.05)
    for c in range(4):
        m = y == c
        mu = z[m].astype(np.float64).mean(0)
        np.testing.assert_allclose(dist[m], np.linalg.norm(z[m] - mu, axis=1), rtol=1e-12)
        assert abs(zs[m].mean()) < 1e-9 and abs(zs[m].std() - 1) < 1e-9
        assert 0.03 <= flags[m].mean() <= 0.07


# ------------------------------------------------------------------------------------------------ N1 classifier
def test_val_transform_matches_reference_golden_and_torchvision_in_process():
    """dataload.py:51-56 restated in numpy: bit-for-bit against the fixture the reference transform produced and
    against torchvision run in-process on other sizes."""
    from oracle import classifier_ref
    g = load_golden("classifier.npz")
    images, _ = classifier_ref.synthetic_eval_set(int(g["n"]), int(g["num_classes"]), seed=int(g["data_seed"]))
    for i in range(4):
        assert np.array_equal(pil_resample.val_transform(images[i]), g["x_head"][i])
    from PIL import Image
    from torchvision import transforms
    tfm = transforms.Compose([transforms.Resize((256, 256)), transforms.CenterCrop(224), transforms.ToTensor(),
                              transforms.Normalize(mean=[0.485, 0.456, 0.406], std=[0.229, 0.224, 0.225])])
    for (h, w), seed in zip([(256, 256), (300, 400), (97, 640), (1024, 300), (224, 224)], range(5)):
        img = np.random.default_rng(seed).integers(0, 256, (h, w, 3), dtype=np.uint8)
        assert np.array_equal(pil_resample.val_transform(img), tfm(Image.fromarray(img)).numpy()), (h, w)


def test_classifier_ref_matches_reference_golden():
    """AnimalClassifier + evaluate_full restated (oracle/classifier_ref.py) against the reference's own outputs."""
    from oracle import classifier_ref
    g = load_golden("classifier.npz")
    n, c, b = int(g["n"]), int(g["num_classes"]), int(g["batch"])
    model = classifier_ref.build_classifier(c, seed=int(g["seed"]))
    images, labels = classifier_ref.synthetic_eval_set(n, c, seed=int(g["data_seed"]))
    batches = classifier_ref.val_batches(images, labels, b)
    import torch
    out = classifier_ref.logits(model, torch.cat([x for x, _ in batches]))
    np.testing.assert_allclose(out, g["logits"], rtol=1e-4, atol=1e-4)
    loss, acc, preds, labs = classifier_ref.evaluate_full(model, batches)
    assert abs(loss - float(g["loss"])) < 1e-4 and acc == float(g["acc"])
    assert np.array_equal(np.asarray(preds), g["preds"]) and np.array_equal(np.asarray(labs), g["labels"])


# ------------------------------------------------------------------------------------------------ N2 Lanczos resize
def _wds_inputs(g):
    from oracle.make_golden import wds_input
    return [wds_input(int(s), int(h), int(w), smooth=(i % 2 == 0)) for i, ((h, w), s) in enumerate(zip(g["sizes"], g["seeds"]))]


def test_wds_lanczos_resize_matches_reference_golden_and_pillow_in_process():
    """data_curation.py:883-913 (resize_and_crop_image, LANCZOS) restated in numpy: bit-for-bit against the fixture
    the reference function produced, and against Pillow run in-process on other sizes."""
    g = load_golden("wds_resize.npz")
    for img, want in zip(_wds_inputs(g), g["crops"]):
        assert np.array_equal(pil_resample.wds_transform_u8(img), want), img.shape
    from PIL import Image
    for (h, w), seed in zip([(260, 333), (90, 120), (224, 230), (512, 380)], range(4)):
        img = np.random.default_rng(seed).integers(0, 256, (h, w, 3), dtype=np.uint8)
        oh, ow = pil_resample.wds_resized_size(h, w)
        ref = np.asarray(Image.fromarray(img).resize((ow, oh), Image.Resampling.LANCZOS))
        assert np.array_equal(pil_resample.resize_bilinear_u8(img, oh, ow, "lanczos"), ref), (h, w)


def test_geometry_of_the_widening_transforms_matches_c_abi_host_helper():
    """irp_preprocess_geometry_ex for the classifier's val_transform and the Lanczos WebDataset resize against the
    oracle's restatement of the reference formulas (sizes, crop offsets, tap bounds), and the Python tap helper."""
    from irp_b200 import _lib
    from irp_b200.stage import taps_for
    rng = np.random.default_rng(3)
    for _ in range(300):
        h, w = int(rng.integers(20, 3000)), int(rng.integers(20, 3000))
        assert _lib.geometry(h, w, _lib.TRANSFORM_VAL_256) == (256, 256, 16, 16, pil_resample.val_max_taps(h, w))
        oh, ow = pil_resample.wds_resized_size(h, w)
        want = (oh, ow, (oh - 224) // 2, (ow - 224) // 2, pil_resample.wds_max_taps(h, w))
        assert _lib.geometry(h, w, _lib.TRANSFORM_WDS_LANCZOS) == want, (h, w)
        for t in (_lib.TRANSFORM_WEIGHTS_DEFAULT, _lib.TRANSFORM_VAL_256, _lib.TRANSFORM_WDS_LANCZOS):
            assert taps_for(h, w, t) == _lib.geometry(h, w, t)[4]


@pytest.mark.parametrize("seed", range(6))
def test_resample_restatements_against_pillow_on_random_sizes(seed):
    """Random sizes (incl. extreme aspect ratios, tiny sides, > 3x downscales): the three transforms' numpy
    restatements are bit-for-bit what Pillow / torchvision produce in-process."""
    from PIL import Image
    from torchvision import transforms
    from torchvision.models import ResNet50_Weights
    rng = np.random.default_rng(1000 + seed)
    h, w = int(rng.integers(8, 900)), int(rng.integers(8, 900))
    img = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
    pil = Image.fromarray(img)
    assert np.array_equal(pil_resample.transform(img), ResNet50_Weights.DEFAULT.transforms()(pil).numpy()), (h, w)
    val = transforms.Compose([transforms.Resize((256, 256)), transforms.CenterCrop(224), transforms.ToTensor(),
                              transforms.Normalize(mean=[0.485, 0.456, 0.406], std=[0.229, 0.224, 0.225])])
    assert np.array_equal(pil_resample.val_transform(img), val(pil).numpy()), (h, w)
    oh, ow = pil_resample.wds_resized_size(h, w)
    r = pil.resize((ow, oh), Image.Resampling.LANCZOS)
    left, top = (ow - 224) // 2, (oh - 224) // 2
    ref = np.asarray(r.crop((left, top, left + 224, top + 224)))
    assert np.array_equal(pil_resample.wds_transform_u8(img), ref), (h, w)


# =============================================================================================== N3 duplicate hash
def test_image_hash_restatement_matches_reference_golden_and_pillow_in_process():
    """oracle/pil_resample.image_hash (bicubic restatement + hashlib.md5) against tests/golden/hash.npz, produced by
    the UNMODIFIED reference compute_image_hash, and against Pillow run here on further sizes."""
    import hashlib
    from PIL import Image
    from oracle.make_golden import wds_input
    g = load_golden("hash.npz")
    imgs = [wds_input(int(s), int(h), int(w), smooth=(i % 2 == 0))
            for i, ((h, w), s) in enumerate(zip(g["sizes"], g["seeds"]))]
    for img, want in zip(imgs, g["hexdigests"]):
        assert pil_resample.image_hash(img) == str(want), img.shape
    near = imgs[4].copy()
    near[150, 200, 1] ^= 0x40
    assert pil_resample.image_hash(near) == str(g["near_hex"]) != str(g["hexdigests"][4])
    rng = np.random.default_rng(5)
    for h, w in [(31, 500), (500, 31), (64, 65), (129, 127), (2000, 90)]:
        a = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        want = hashlib.md5(Image.fromarray(a).resize((64, 64)).convert("RGB").tobytes()).hexdigest()
        assert pil_resample.image_hash(a) == want
        assert _lib_geometry_taps(h, w) == pil_resample.hash_max_taps(h, w)


def _lib_geometry_taps(h, w):
    from irp_b200 import _lib
    return _lib.geometry(h, w, _lib.TRANSFORM_HASH_64)[4]


# ------------------------------------------------------------------ N4 UMAP graph construction (parity unpinned)
def test_umap_graph_ref_neighbours_match_sklearn_bruteforce_and_the_defining_equations():
    """umap-learn is absent: the restatement is pinned to scikit-learn's exact neighbours and to the equations of
    umap_.py smooth_knn_dist / compute_membership_strengths / fuzzy_simplicial_set."""
    from sklearn.neighbors import NearestNeighbors
    from oracle import umap_graph_ref as ug
    rng = np.random.default_rng(0)
    x = rng.normal(size=(600, 20)).astype(np.float32)
    x[5] = x[6]  # an exact duplicate: rho must skip the zero distance
    k = 15
    idx, d = ug.nearest_neighbors(x, k)
    dd, ii = NearestNeighbors(n_neighbors=k, algorithm="brute").fit(x.astype(np.float64)).kneighbors(
        x.astype(np.float64))
    assert np.abs(d - dd).max() <= 1e-6
    for i in range(600):  # equal up to the order of exactly tied distances
        assert set(idx[i].tolist()) == set(ii[i].tolist()) or np.unique(dd[i]).size < k
    sig, rho = ug.smooth_knn_dist(d, float(k))
    first_pos = np.array([row[row > 0][0] for row in d])
    assert np.array_equal(rho, first_pos)
    psum = np.exp(-np.maximum(d[:, 1:] - rho[:, None], 0) / sig[:, None]).sum(1)
    assert np.abs(psum - np.log2(k)).max() < 1e-4
    rows, cols, vals = ug.compute_membership_strengths(idx, d, sig, rho)
    v = vals.reshape(600, k)
    assert np.all(v[idx == np.arange(600)[:, None]] == 0) and np.all((v >= 0) & (v <= 1))
    nearest = np.array([np.flatnonzero(row > 0)[0] for row in d])
    assert np.all(v[np.arange(600), nearest] == 1.0)  # local_connectivity = 1: the nearest neighbour has weight 1
    g, _, _ = ug.fuzzy_simplicial_set(x, k)
    g = g.tocsr()
    assert abs(g - g.T).max() == 0 and g.data.min() > 0 and g.data.max() <= 1.0
    a = np.zeros((600, 600), np.float64)
    keep = cols != rows
    a[rows[keep], cols[keep]] = vals[keep]
    assert np.abs(g.toarray() - (a + a.T - a * a.T)).max() <= 1e-6


def test_create_embeddings_hands_device_knn_to_umap(monkeypatch):
    """The drop-in passes UMAP's own precomputed_knn parameter (host logic only: the device call is stubbed)."""
    import sys
    import types
    from functions import data_curation as dc
    from irp_b200 import umap_graph
    calls = {}
    monkeypatch.setattr(umap_graph, "nearest_neighbors",
                        lambda x, k, device=None: calls.setdefault("knn", (x.shape, k)) and ("IDX", "DIST"))
    z = np.zeros((40, 5), np.float32)
    p = dc._with_device_knn({"n_components": 2, "random_state": 42}, z)
    assert p["precomputed_knn"] == ("IDX", "DIST") and calls["knn"] == ((40, 5), 15)
    assert "precomputed_knn" not in dc._with_device_knn({"metric": "cosine"}, z)
    assert dc._with_device_knn({"precomputed_knn": (1, 2)}, z)["precomputed_knn"] == (1, 2)
    assert "precomputed_knn" not in dc._with_device_knn({"n_neighbors": 50}, z)  # fewer rows than neighbours
