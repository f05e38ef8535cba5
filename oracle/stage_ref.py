"""ORACLE (test infrastructure, never on the product path): the reference's own torch-CPU route for the stage.

Restates /root/reference/functions/data_curation.py:654-728 with the same third-party calls the reference makes
(torchvision ResNet-50 + transform preset, PIL decode, sklearn PCA / LocalOutlierFactor), with the two changes
the survey requires to make it runnable and reproducible here:

* D5: ``initialize_model`` builds ``resnet50(weights=None)`` under a fixed seed instead of downloading
  ``ResNet50_Weights.DEFAULT`` (:656-657); the Sequential(*children[:-1]).eval() wrapping (:658-659) and the
  transform preset are unchanged;
* D4: PCA uses ``svd_solver="full"`` on float64 (the exact solver) because the reference's default randomized,
  unseeded solver does not reproduce itself (:700-701).

``bench.py --impl reference`` and the ``cpu_baseline`` leg time these functions on the host cores; tests use them
as the checker.  oracle/make_golden.py pins them against the UNMODIFIED reference module imported from
/root/reference (available in the build container only) and stores the outputs under tests/golden/.
"""
from __future__ import annotations

import os

import numpy as np
import torch

EMBED_DIM = 2048


def initialize_model(device="cpu", seed=1234):
    """data_curation.py:654-659 with random-init weights (D5): (model, transform)."""
    from torchvision import models
    from torchvision.models import ResNet50_Weights

    torch.manual_seed(seed)
    model = models.resnet50(weights=None)
    model = torch.nn.Sequential(*list(model.children())[:-1])
    return model.to(device).eval(), ResNet50_Weights.DEFAULT.transforms()


def full_resnet50(seed=1234):
    """The un-truncated torchvision module under the same seed (source of the weights the CUDA trunk loads)."""
    from torchvision import models

    torch.manual_seed(seed)
    return models.resnet50(weights=None).eval()


def process_image_directory(root_dir, device="cpu", transform=None, batch_size=32, seed=1234):
    """data_curation.py:661-684, literally: one image at a time, skip-and-print on failure."""
    from PIL import Image

    model, tfm = initialize_model(device, seed)
    transform = transform or tfm
    features, labels, paths = [], [], []
    for class_name in os.listdir(root_dir):
        class_dir = os.path.join(root_dir, class_name)
        if not os.path.isdir(class_dir):
            continue
        for img_name in os.listdir(class_dir):
            img_path = os.path.join(class_dir, img_name)
            try:
                img = Image.open(img_path).convert("RGB")
                img_tensor = transform(img).unsqueeze(0).to(device)
                with torch.no_grad():
                    feat = model(img_tensor).cpu().squeeze().numpy()
                features.append(feat)
                labels.append(class_name)
                paths.append(img_path)
            except Exception as e:  # noqa: BLE001
                print(f"Skipped {img_path}: {str(e)}")
    return np.array(features), np.array(labels), np.array(paths)


def embed_arrays(images, model=None, transform=None, batch_size=32, seed=1234, device="cpu"):
    """Same arithmetic as process_image_directory on in-memory HWC uint8 arrays, batched (config 1 wording:
    'batch 32').  Returns float32 [n,2048]."""
    from PIL import Image

    if model is None:
        model, transform = initialize_model(device, seed)
    out = []
    with torch.no_grad():
        for s in range(0, len(images), batch_size):
            x = torch.stack([transform(Image.fromarray(im)) for im in images[s:s + batch_size]]).to(device)
            out.append(model(x).flatten(1).cpu().numpy())
    return np.concatenate(out, 0) if out else np.zeros((0, EMBED_DIM), np.float32)


def pca_exact(features, pca_components=50):
    """PCA half of create_embeddings (:700-701) with the exact solver (D4): (features_pca float64, pca)."""
    from sklearn.decomposition import PCA

    pca = PCA(n_components=pca_components, svd_solver="full")
    z = pca.fit_transform(np.asarray(features, np.float64))
    return z, pca


def pca_default(features, pca_components=50):
    """The reference's literal call (:700-701): default solver, unseeded (behavioural reference + timing only)."""
    from sklearn.decomposition import PCA

    pca = PCA(n_components=pca_components)
    return pca.fit_transform(features), pca


def detect_outliers(embedding, labels, class_n_neighbors=30, class_contamination=0.05, global_n_neighbors=75,
                    global_contamination=0.03):
    """data_curation.py:709-728 with the same sklearn calls."""
    from sklearn.neighbors import LocalOutlierFactor
    from sklearn.preprocessing import LabelEncoder

    le = LabelEncoder()
    y_numeric = le.fit_transform(labels)
    class_outliers = np.zeros(len(labels), dtype=bool)
    for class_id in np.unique(y_numeric):
        mask = y_numeric == class_id
        lof = LocalOutlierFactor(n_neighbors=class_n_neighbors, contamination=class_contamination)
        class_outliers[mask] = lof.fit_predict(embedding[mask]) == -1
    global_lof = LocalOutlierFactor(n_neighbors=global_n_neighbors, contamination=global_contamination)
    global_outliers = global_lof.fit_predict(embedding) == -1
    return class_outliers, global_outliers


def lof_band(embedding, n_neighbors, contamination, rel=1e-3):
    """(flags, near_threshold mask): rows whose score lies within rel*|offset| of the threshold may legitimately
    flip under a different summation order (north_star: 'identical except within 1e-3 of the threshold')."""
    from sklearn.neighbors import LocalOutlierFactor

    lof = LocalOutlierFactor(n_neighbors=n_neighbors, contamination=contamination)
    flags = lof.fit_predict(embedding) == -1
    near = np.abs(lof.negative_outlier_factor_ - lof.offset_) <= rel * abs(lof.offset_)
    return flags, near
