"""Seeded synthetic inputs shared by the oracle, the tests and bench.py.

Test/bench infrastructure only.  The reference ships no data; its recorded dataset statistics
(/root/reference/dataset_cleaning_report.txt:9-31: 26 003 images, 10 classes with 1 433..4 849 images;
/root/reference/dataset_analysis_report.txt:25-29,46-51: mean 320x253 px, 45.7 % with a side < 224) size the
"Animals-10-sized" configuration.
"""
from __future__ import annotations

import numpy as np

# class sizes of the cleaned Animals-10 set (dataset_cleaning_report.txt:22-31), used as proportions
ANIMALS10_CLASS_SIZES = [4849, 2613, 1433, 2102, 3086, 1663, 1855, 1812, 4795, 1795]
CLASS_NAMES = ["butterfly", "cat", "chicken", "cow", "dog", "elephant", "horse", "sheep", "spider", "squirrel"]


def smooth_image(rng: np.random.Generator, h: int, w: int, class_id: int) -> np.ndarray:
    """Smooth class-tinted uint8 HWC image: 8x8 random base upsampled bicubically + noise + class offset."""
    from PIL import Image

    base = rng.integers(0, 256, (8, 8, 3), dtype=np.uint8)
    up = np.asarray(Image.fromarray(base).resize((w, h), Image.BICUBIC)).astype(np.int16)
    noise = rng.integers(-20, 20, (h, w, 3), dtype=np.int16)
    return np.clip(up + noise + 5 * class_id, 0, 255).astype(np.uint8)


def config1_images(n: int = 1024, n_classes: int = 10, seed: int = 0, size: int = 224):
    """Config 1: n 224x224 images, classes round-robin.  Returns (list of HWC uint8, list of class names)."""
    rng = np.random.default_rng(seed)
    imgs, labels = [], []
    for i in range(n):
        c = i % n_classes
        imgs.append(smooth_image(rng, size, size, c))
        labels.append(f"class{c}")
    return imgs, labels


def mixed_resolution_sizes(n: int, seed: int = 0) -> np.ndarray:
    """(n, 2) int32 heights/widths around 300x400: H in U[200,400], W in U[250,500]; ~45 % get a side < 224
    (upscaling path) and ~0.2 % are large (>= 1500 px) to exercise many-tap antialiasing."""
    rng = np.random.default_rng(seed)
    h = rng.integers(200, 401, n)
    w = rng.integers(250, 501, n)
    small = rng.random(n) < 0.42
    h = np.where(small, rng.integers(120, 224, n), h)
    portrait = rng.random(n) < 0.235
    h2 = np.where(portrait, w, h)
    w2 = np.where(portrait, h, w)
    big = rng.random(n) < 0.002
    h2 = np.where(big, rng.integers(1500, 2600, n), h2)
    w2 = np.where(big, rng.integers(1500, 2600, n), w2)
    return np.stack([h2, w2], axis=1).astype(np.int32)


def class_assignment(n: int, seed: int = 0) -> np.ndarray:
    """Class ids with the Animals-10 proportions, shuffled (int32 [n])."""
    rng = np.random.default_rng(seed + 1)
    sizes = np.array(ANIMALS10_CLASS_SIZES, np.float64)
    counts = np.floor(sizes / sizes.sum() * n).astype(np.int64)
    counts[0] += n - counts.sum()
    ids = np.repeat(np.arange(len(sizes)), counts)
    rng.shuffle(ids)
    return ids.astype(np.int32)


def embedding_like(n: int, d: int = 2048, seed: int = 0) -> np.ndarray:
    """float32 [n,d] with the shape of random-init ResNet-50 embeddings: dominant shared component, fast-decaying
    head, slowly decaying tail, non-negative with exact zeros."""
    rng = np.random.default_rng(seed)
    r = min(256, d)
    basis = np.linalg.qr(rng.standard_normal((d, r)))[0]
    sv = np.concatenate([[150.0, 60.0, 30.0], 12.0 * np.arange(1, r - 2) ** -0.6])[:r]
    x = (rng.standard_normal((n, r)) * sv) @ basis.T + 0.05 * rng.standard_normal((n, d)) + 3.0
    return np.maximum(x, 0).astype(np.float32)


def clustered_points(n: int, d: int, n_classes: int, seed: int = 0):
    """float32 [n,d] Gaussian clusters with varying spread + int class ids (LOF test input)."""
    rng = np.random.default_rng(seed)
    y = rng.integers(0, n_classes, n)
    centers = rng.standard_normal((n_classes, d)) * 4
    z = centers[y] + rng.standard_normal((n, d)) * rng.uniform(0.5, 2.0, (n, 1))
    return z.astype(np.float32), y.astype(np.int64)
