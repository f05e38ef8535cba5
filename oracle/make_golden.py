"""Generates tests/golden/*.npz by running the UNMODIFIED reference (imported from /root/reference) on seeded inputs.

Run in the build container only (the GPU box has no /root/reference):  python oracle/make_golden.py
The fixtures pin the oracle restatements (tests/test_oracle.py) and the CUDA path (tests/test_gpu_parity.py).

Shims: empty stub modules for the plotting / notebook / umap / webdataset
imports at the top of data_curation.py (none is touched by the functions exercised here), and
``data_curation.initialize_model`` replaced by the seeded random-init builder of oracle/stage_ref.py (D5) -- it
is looked up as a module global at data_curation.py:663, so process_image_directory itself runs unmodified.
"""
from __future__ import annotations

import importlib.machinery
import os
import sys
import tempfile
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")
REFERENCE = "/root/reference"


def import_reference():
    def stub(name, attrs=()):
        m = types.ModuleType(name)
        m.__spec__ = importlib.machinery.ModuleSpec(name, None)
        for a in attrs:
            setattr(m, a, type(a, (), {}))
        sys.modules[name] = m
        return m

    mpl = stub("matplotlib")
    mpl.pyplot = stub("matplotlib.pyplot")
    mpl.figure = stub("matplotlib.figure", ["Figure"])
    ip = stub("IPython")
    ipd = stub("IPython.display", ["Markdown"])
    ipd.display = lambda *a, **k: None
    ip.display = ipd
    um = stub("umap")
    um.umap_ = stub("umap.umap_", ["UMAP"])
    stub("webdataset")
    sys.path.insert(0, REFERENCE)
    import functions.data_curation as dc  # the reference module, unmodified

    assert dc.__file__.startswith(REFERENCE), dc.__file__
    return dc


def golden_preprocess(dc):
    """Reference transform (ResNet50_Weights.DEFAULT.transforms(), data_curation.py:659,675) on assorted sizes."""
    import hashlib

    from PIL import Image
    from torchvision.models import ResNet50_Weights

    tfm = ResNet50_Weights.DEFAULT.transforms()
    rng = np.random.default_rng(2024)
    sizes = [(224, 224), (300, 400), (400, 300), (150, 200), (57, 60), (700, 500), (233, 232), (231, 500)]
    seeds, crops, hashes = [], [], []
    for i, (h, w) in enumerate(sizes):
        seed = 1000 + i
        img = np.random.default_rng(seed).integers(0, 256, (h, w, 3), dtype=np.uint8)
        out = tfm(Image.fromarray(img)).numpy()  # float32 [3,224,224]
        # the uint8 pixels behind it: invert the normalisation exactly via a lookup over the 256 levels
        mean = np.array([0.485, 0.456, 0.406], np.float32)
        std = np.array([0.229, 0.224, 0.225], np.float32)
        levels = ((np.arange(256, dtype=np.float32) / np.float32(255.0))[:, None] - mean) / std  # [256,3]
        u8 = np.empty((224, 224, 3), np.uint8)
        for c in range(3):
            idx = np.searchsorted(levels[:, c], out[c])
            idx = np.clip(idx, 0, 255)
            lo = np.clip(idx - 1, 0, 255)
            pick = np.where(np.abs(levels[lo, c] - out[c]) <= np.abs(levels[idx, c] - out[c]), lo, idx)
            assert np.array_equal(levels[pick, c], out[c]), "normalisation is not a pure per-level map"
            u8[:, :, c] = pick
        seeds.append(seed)
        crops.append(u8)
        hashes.append(hashlib.sha256(out.tobytes()).hexdigest())
    del rng
    np.savez_compressed(os.path.join(GOLDEN, "preprocess.npz"), sizes=np.array(sizes, np.int32),
                        seeds=np.array(seeds, np.int64), crops=np.stack(crops), float32_sha256=np.array(hashes))
    print("preprocess.npz:", len(sizes), "images")


def golden_embeddings(dc):
    """Unmodified process_image_directory on a small seeded PNG tree (random-init weights, seed 1234)."""
    from PIL import Image

    from oracle import stage_ref, synth

    dc.initialize_model = lambda device: stage_ref.initialize_model(device, seed=1234)
    rng = np.random.default_rng(7)
    sizes = [(96, 128), (128, 96), (224, 224), (180, 260), (120, 120), (260, 330)]
    images, names = [], []
    with tempfile.TemporaryDirectory() as td:
        for i in range(12):
            c = i % 4
            h, w = sizes[i % len(sizes)]
            img = synth.smooth_image(rng, h, w, c)
            d = os.path.join(td, f"class{c}")
            os.makedirs(d, exist_ok=True)
            name = f"class{c}/img{i:03d}.png"
            Image.fromarray(img).save(os.path.join(td, name))
            images.append(img)
            names.append(name)
        # a file the reference must skip (:681-682) and a stray non-directory entry (:668-669)
        with open(os.path.join(td, "class0", "broken.png"), "wb") as f:
            f.write(b"not an image")
        with open(os.path.join(td, "README.txt"), "w") as f:
            f.write("stray file")
        tfm = stage_ref.initialize_model("cpu")[1]
        feats, labels, paths = dc.process_image_directory(td, "cpu", tfm)
        rel = [os.path.relpath(p, td) for p in paths]
    order = [rel.index(n) for n in names]  # align to our image order (os.listdir order is fs-dependent, H6)
    assert feats.shape == (12, 2048) and len(rel) == 12
    flat = np.concatenate([im.reshape(-1) for im in images])
    np.savez_compressed(os.path.join(GOLDEN, "embeddings.npz"), pixels=flat,
                        hw=np.array([im.shape[:2] for im in images], np.int32), names=np.array(names),
                        labels=labels[order], features=feats[order].astype(np.float32), seed=np.int64(1234))
    print("embeddings.npz:", feats.shape, "skipped files handled:", len(rel) == 12)


def golden_pca(dc):
    """sklearn PCA exactly as the reference constructs it, but svd_solver='full' on float64 (D4)."""
    from sklearn.decomposition import PCA

    from oracle import synth

    x = synth.embedding_like(600, 2048, seed=11)
    k = 20
    pca = PCA(n_components=k, svd_solver="full")
    z = pca.fit_transform(x.astype(np.float64))
    np.savez_compressed(os.path.join(GOLDEN, "pca.npz"), n=np.int64(600), d=np.int64(2048), seed=np.int64(11),
                        k=np.int64(k), mean=pca.mean_, components=pca.components_,
                        explained_variance=pca.explained_variance_,
                        explained_variance_ratio=pca.explained_variance_ratio_,
                        singular_values=pca.singular_values_, noise_variance=np.float64(pca.noise_variance_),
                        z_head=z[:64])
    print("pca.npz: k =", k, "lambda_1 =", pca.explained_variance_[0], "lambda_k =", pca.explained_variance_[-1])


def golden_lof(dc):
    """Unmodified detect_outliers (data_curation.py:709-728) on seeded clustered points."""
    import warnings

    from sklearn.neighbors import LocalOutlierFactor

    from oracle import synth

    z, y = synth.clustered_points(900, 50, 10, seed=5)
    labels = np.array([f"cls{c:02d}" for c in y])
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        cls_out, glob_out = dc.detect_outliers(z, labels)
        lof = LocalOutlierFactor(n_neighbors=75, contamination=0.03).fit(z)
    # 2-D variant (what the reference feeds in production: UMAP output) with tiny classes that clip k
    z2, y2 = synth.clustered_points(300, 2, 12, seed=6)
    labels2 = np.array([f"c{c:02d}" for c in y2])
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        cls2, glob2 = dc.detect_outliers(z2, labels2)
    np.savez_compressed(os.path.join(GOLDEN, "lof.npz"), n=np.int64(900), d=np.int64(50), classes=np.int64(10),
                        seed=np.int64(5), class_outliers=cls_out, global_outliers=glob_out,
                        global_scores=lof.negative_outlier_factor_.astype(np.float64),
                        global_offset=np.float64(lof.offset_), n2=np.int64(300), d2=np.int64(2),
                        classes2=np.int64(12), seed2=np.int64(6), class_outliers2=cls2, global_outliers2=glob2)
    print("lof.npz: flagged", int(cls_out.sum()), int(glob_out.sum()), "| 2-D:", int(cls2.sum()), int(glob2.sum()))


def wds_input(seed, h, w, smooth):
    """Test image for the Lanczos fixtures: uniform noise, or blocky smooth content (8x8 cells) plus mild noise --
    numpy only, so the tests can rebuild the inputs without depending on any resampling code."""
    rng = np.random.default_rng(seed)
    if not smooth:
        return rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
    cells = rng.integers(0, 256, ((h + 7) // 8, (w + 7) // 8, 3)).astype(np.int32)
    img = np.repeat(np.repeat(cells, 8, axis=0), 8, axis=1)[:h, :w] + rng.integers(-12, 13, (h, w, 3))
    return np.clip(img, 0, 255).astype(np.uint8)


def golden_wds_resize(dc):
    """the reference's resize_and_crop_image (data_curation.py:883-913), unmodified, on RGB inputs
    of assorted sizes (up- and downscales, both orientations, already-224 sides)."""
    from PIL import Image

    sizes = [(224, 224), (300, 400), (400, 300), (150, 200), (57, 60), (700, 500), (225, 224), (231, 500),
             (224, 600), (1000, 1300)]
    seeds, outs = [], []
    for i, (h, w) in enumerate(sizes):
        seed = 3000 + i
        img = wds_input(seed, h, w, smooth=(i % 2 == 0))
        out = np.asarray(dc.resize_and_crop_image(Image.fromarray(img)))
        assert out.shape == (224, 224, 3) and out.dtype == np.uint8
        seeds.append(seed)
        outs.append(out)
    np.savez_compressed(os.path.join(GOLDEN, "wds_resize.npz"), sizes=np.array(sizes, np.int32),
                        seeds=np.array(seeds, np.int64), crops=np.stack(outs))
    print("wds_resize.npz:", len(sizes), "images")


def golden_hash(dc):
    """the reference's compute_image_hash (data_curation.py:283-292), unmodified, on RGB inputs of
    assorted sizes (the dataset's smallest 60x57, up- and downscales, one that needs 40+ taps), plus an exact
    duplicate and a one-pixel-off near-duplicate (the pair the dataset scan at :394-399 must tell apart)."""
    from PIL import Image

    sizes = [(57, 60), (64, 64), (224, 224), (253, 320), (300, 400), (400, 300), (700, 500), (64, 900),
             (1000, 1300), (1500, 1400)]
    seeds, hexes = [], []
    for i, (h, w) in enumerate(sizes):
        seed = 5000 + i
        img = wds_input(seed, h, w, smooth=(i % 2 == 0))
        hexes.append(dc.compute_image_hash(Image.fromarray(img)))
        seeds.append(seed)
    dup = wds_input(5004, 300, 400, smooth=True)           # same pixels as entry 4
    near = dup.copy()
    near[150, 200, 1] ^= 0x40                               # one channel of one pixel changed
    extra = [dc.compute_image_hash(Image.fromarray(dup)), dc.compute_image_hash(Image.fromarray(near))]
    assert extra[0] == hexes[4] and extra[1] != hexes[4]
    np.savez_compressed(os.path.join(GOLDEN, "hash.npz"), sizes=np.array(sizes, np.int32),
                        seeds=np.array(seeds, np.int64), hexdigests=np.array(hexes), near_hex=np.array(extra[1]))
    print("hash.npz:", len(sizes), "images", hexes[:2])


def golden_classifier():
    """the reference's val_transform, AnimalClassifier and evaluate_full, unmodified except that
    ``functions.model.resnet50`` is patched to build the random-init network (no download; the seed is set right
    before the constructor runs so the parameters equal oracle/classifier_ref.build_classifier(seed))."""
    import importlib.machinery as mach
    import types as _types

    import torch
    from PIL import Image

    for name in ("mlflow",):
        m = _types.ModuleType(name)
        m.__spec__ = mach.ModuleSpec(name, None)
        sys.modules.setdefault(name, m)
    import torchvision.models as tvm

    import functions.dataload as dataload  # reference, unmodified (webdataset stubbed by import_reference)
    import functions.model as ref_model
    import functions.train as ref_train
    from oracle import classifier_ref

    assert ref_model.__file__.startswith(REFERENCE) and ref_train.__file__.startswith(REFERENCE)
    ref_model.resnet50 = lambda weights=None: tvm.resnet50(weights=None)
    seed, num_classes, n = 1234, 10, 48
    torch.manual_seed(seed)
    model = ref_model.AnimalClassifier(num_classes=num_classes).eval()
    images, labels = classifier_ref.synthetic_eval_set(n, num_classes, seed=0)
    _, val_transform = dataload.get_transforms("medium")
    xs = torch.stack([val_transform(Image.fromarray(im)) for im in images])
    batches = [(xs[s:s + 16], torch.from_numpy(labels[s:s + 16])) for s in range(0, n, 16)]
    with torch.no_grad():
        logits = model(xs).numpy()
    ref_train.DEVICE = "cpu"
    loss, acc, preds, labs = ref_train.evaluate_full(model, batches, torch.nn.CrossEntropyLoss(),
                                                     disable_progress=True)
    np.savez_compressed(os.path.join(GOLDEN, "classifier.npz"), seed=np.int64(seed), num_classes=np.int64(num_classes),
                        n=np.int64(n), data_seed=np.int64(0), batch=np.int64(16),
                        x_head=xs[:4].numpy(), logits=logits, loss=np.float64(loss), acc=np.float64(acc),
                        preds=np.asarray(preds, np.int64), labels=np.asarray(labs, np.int64))
    print("classifier.npz: loss %.6f acc %.2f" % (loss, acc), "logit range", logits.min(), logits.max())


if __name__ == "__main__":
    os.makedirs(GOLDEN, exist_ok=True)
    dc = import_reference()
    if len(sys.argv) > 1 and sys.argv[1] == "classifier":
        golden_classifier()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "wds":
        golden_wds_resize(dc)
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "hash":
        golden_hash(dc)
        sys.exit(0)
    golden_preprocess(dc)
    golden_embeddings(dc)
    golden_pca(dc)
    golden_lof(dc)
    golden_classifier()
    golden_wds_resize(dc)
    golden_hash(dc)
