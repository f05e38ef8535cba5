"""Drop-in mirror of the reference's outlier-analysis entry points on the B200 kernels.

Same names, argument meaning, return types and error behaviour as
/root/reference/functions/data_curation.py:654-743 (``initialize_model``, ``process_image_directory``,
``create_embeddings``, ``detect_outliers``, ``create_results_dataframe``); the work is done by libirp_b200.so
through the ``irp_b200`` torch custom ops.  Put ``image-recognition-pipeline_b200`` on ``sys.path`` and
``from functions import data_curation`` keeps working in the notebook.

Deliberate, documented differences:
* ``batch_size`` is honoured (the reference accepts it and then runs one image at a time, :661,:675);
* the PCA is the exact covariance route (``svd_solver='full'`` semantics) instead of sklearn's unseeded
  randomized solver, so results are reproducible (D4);
* ``initialize_model`` accepts ``weights=`` so random-init / local state dicts work without network access (D5);
* UMAP stays host-side and optional: ``create_embeddings`` raises ImportError if umap-learn is absent, exactly where
  the reference module would fail at import time; ``create_pca_embeddings`` exposes the accelerated half.
"""
from __future__ import annotations

import os
import warnings

import numpy as np
import torch

from irp_b200 import _lib, ops
from irp_b200.stage import OutlierStage, ResNet50Trunk, pack_images

_DEFAULT_MAX_BATCH = 256
_trunk_cache = {}


class B200Transform:
    """`weights.transforms()` replacement: resize 232 (antialiased bilinear) -> center crop 224 -> /255 ->
    normalise, on the GPU (irp_preprocess).  Callable on a PIL image like the reference's transform (returns the
    [3,224,224] tensor); `process_image_directory` recognises it and runs whole batches through the fused kernel."""

    def __init__(self, device):
        self.device = torch.device(device)

    def __call__(self, img) -> torch.Tensor:
        arr = np.asarray(img.convert("RGB") if hasattr(img, "convert") else img, dtype=np.uint8)
        packed = pack_images([arr]).to(self.device)
        out = ops.preprocess(packed.pixels, packed.offsets, packed.hw, packed.max_taps, _lib.LAYOUT_NCHW)
        return out[0].float()

    def __repr__(self):
        return "B200Transform(resize=232, crop=224, mean=[0.485,0.456,0.406], std=[0.229,0.224,0.225])"


class B200ResNet50(torch.nn.Module):
    """`Sequential(*resnet50.children()[:-1]).eval()` replacement: [B,3,224,224] -> [B,2048,1,1]."""

    def __init__(self, trunk: ResNet50Trunk):
        super().__init__()
        self.trunk = trunk

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return self.trunk.embed_nchw(x).view(-1, _lib.EMBED_DIM, 1, 1)

    def to(self, *args, **kwargs):  # the trunk is pinned to its device
        return self

    def eval(self):
        return self


def _build_torch_resnet50(weights, seed):
    from torchvision import models
    from torchvision.models import ResNet50_Weights

    if weights == "DEFAULT":
        return models.resnet50(weights=ResNet50_Weights.DEFAULT)  # same download as the reference (:656-657)
    if weights is None:
        if seed is not None:
            torch.manual_seed(seed)
        return models.resnet50(weights=None)
    if isinstance(weights, torch.nn.Module):
        return weights
    m = models.resnet50(weights=None)
    m.load_state_dict(weights)
    return m


def initialize_model(device, weights="DEFAULT", seed=None, max_batch=_DEFAULT_MAX_BATCH):
    """Initialize ResNet50 model with pretrained weights  (reference: data_curation.py:654-659).

    Returns (model, transform) like the reference.  `weights`: "DEFAULT" (ImageNet, downloads), None (random init
    under `seed`), a state dict or a torchvision ResNet-50 module."""
    device = torch.device(device)
    tm = _build_torch_resnet50(weights, seed).eval()
    trunk = ResNet50Trunk(tm, device, max_batch=max_batch)
    return B200ResNet50(trunk), B200Transform(device)


def _stage_for(model, batch_size):
    trunk = model.trunk
    return OutlierStage(trunk, batch_size=min(batch_size, trunk.max_batch))


def process_image_directory(root_dir, device, transform, batch_size=32, model=None):
    """Process directory and extract features  (reference: data_curation.py:661-684).

    Walks `root_dir/<class>/<image>` in os.listdir order, decodes with PIL on the host, and embeds `batch_size`
    images per launch.  Images that fail to decode/convert are reported as ``Skipped <path>: <error>`` and left
    out of all three returned arrays, as in the reference (:681-682)."""
    from PIL import Image

    if model is None:
        model, _ = initialize_model(device)  # the reference rebuilds the model here as well (:663)
    stage = _stage_for(model, batch_size)
    features, labels, paths = [], [], []
    pend_imgs, pend_labels, pend_paths = [], [], []
    fused = isinstance(transform, B200Transform)

    def flush():
        # Skip-and-print is a PER-IMAGE contract (:681-682) and covers what can fail per image: decode, convert and
        # the host-side transform, all handled in the loop below.  A failure of the batched device path (IrpError,
        # a CUDA error) is not an image problem: it propagates instead of silently dropping up to a whole batch.
        if not pend_imgs:
            return
        if fused:
            feats = stage.embed_packed(pack_images(pend_imgs), from_host=True)
        else:
            feats = model.trunk.embed_nchw(torch.stack(pend_imgs))
        features.append(np.array(stage.to_host(feats, "batch").numpy()))  # pinned staging buffer, then an owned copy
        labels.extend(pend_labels)
        paths.extend(pend_paths)
        pend_imgs.clear()
        pend_labels.clear()
        pend_paths.clear()

    for class_name in os.listdir(root_dir):
        class_dir = os.path.join(root_dir, class_name)
        if not os.path.isdir(class_dir):
            continue
        for img_name in os.listdir(class_dir):
            img_path = os.path.join(class_dir, img_name)
            try:
                img = Image.open(img_path).convert('RGB')
                item = np.asarray(img, dtype=np.uint8) if fused else transform(img)
                if fused and (item.ndim != 3 or item.shape[2] != 3 or item.shape[0] < 1 or item.shape[1] < 1):
                    raise ValueError(f"unexpected decoded shape {item.shape}")
                pend_imgs.append(item)
                pend_labels.append(class_name)
                pend_paths.append(img_path)
            except Exception as e:
                print(f"Skipped {img_path}: {str(e)}")
            if len(pend_imgs) >= stage.batch_size:
                flush()
    flush()
    feats = np.concatenate(features, 0) if features else np.zeros((0,), np.float32)
    return feats, np.array(labels), np.array(paths)


def _fit_pca(features, pca_components, device=None):
    """GPU PCA -> (features_pca float32 [n,k], fitted sklearn PCA object)."""
    from sklearn.decomposition import PCA

    device = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
    x = torch.from_numpy(np.ascontiguousarray(features, dtype=np.float32)).to(device)
    n, d = x.shape
    if not 1 <= pca_components <= min(n, d):
        raise ValueError(f"n_components={pca_components!r} must be between 0 and min(n_samples, n_features)="
                         f"{min(n, d)!r} with svd_solver='full'")
    shift = x[: min(256, n)].mean(0).contiguous()
    acc = torch.zeros(1 + d + d * d, dtype=torch.float64, device=device)
    count, total, scatter = acc[:1], acc[1:1 + d], acc[1 + d:].view(d, d)
    ops.cov_accumulate(x, shift, count, total, scatter)
    mean, comps, evals = ops.pca_fit(count, total, scatter, shift, pca_components)
    z = ops.pca_transform(x, mean, comps)
    ev = evals[:pca_components].cpu().numpy()
    total_var = float(evals[pca_components].item())
    pca = PCA(n_components=pca_components, svd_solver="full")
    dt = np.float32 if np.asarray(features).dtype == np.float32 else np.float64
    pca.mean_ = mean.cpu().numpy().astype(dt)
    pca.components_ = comps.cpu().numpy().astype(dt)
    pca.explained_variance_ = ev.astype(dt)
    pca.explained_variance_ratio_ = (ev / total_var).astype(dt)
    pca.singular_values_ = np.sqrt(ev * (n - 1)).astype(dt)
    pca.n_components_ = pca_components
    pca.n_samples_ = n
    pca.n_features_in_ = d
    m = min(n, d)
    pca.noise_variance_ = float((total_var - ev.sum()) / (m - pca_components)) if pca_components < m else 0.0
    pca._fit_svd_solver = "full"
    return z.cpu().numpy(), pca


def create_pca_embeddings(features, labels, pca_components=50):
    """The accelerated half of create_embeddings: LabelEncoder + PCA (reference: data_curation.py:696-701)."""
    from sklearn.preprocessing import LabelEncoder

    le = LabelEncoder()
    le.fit_transform(labels)
    features_pca, pca = _fit_pca(features, pca_components)
    return features_pca, le, pca


def create_embeddings(features, labels, pca_components=50, umap_params=None):
    """Create supervised UMAP embeddings  (reference: data_curation.py:686-707).

    PCA runs on the GPU; UMAP stays host-side (north_star) and consumes the new PCA output."""
    umap_params = umap_params or {
        'n_components': 2,
        'target_metric': 'categorical',
        'target_weight': 0.5,
        'random_state': 42,
        'n_jobs': -1
    }
    try:
        import umap.umap_ as umap
    except ImportError as e:  # the reference fails at module import (data_curation.py:21)
        raise ImportError("umap-learn is required for create_embeddings (host-side UMAP); "
                          "use create_pca_embeddings for the PCA half") from e
    from sklearn.preprocessing import LabelEncoder

    le = LabelEncoder()
    y_numeric = le.fit_transform(labels)
    features_pca, pca = _fit_pca(features, pca_components)
    reducer = umap.UMAP(**_with_device_knn(umap_params, features_pca))
    embedding = reducer.fit_transform(features_pca, y=y_numeric)
    return embedding, le, pca, reducer


def _with_device_knn(umap_params, features_pca):
    """UMAP's neighbour search (umap_.py nearest_neighbors) runs on the device -- exactly,
    where umap-learn switches to approximate NN-descent above 4 096 samples -- and is handed over through UMAP's own
    `precomputed_knn=(knn_indices, knn_dists)` parameter; everything else of UMAP stays host-side.  Left alone when
    the caller brings a `precomputed_knn`, a non-Euclidean or precomputed metric, or fewer rows than neighbours."""
    params = dict(umap_params)
    k = int(params.get('n_neighbors', 15))
    if ('precomputed_knn' in params or params.get('metric', 'euclidean') != 'euclidean'
            or features_pca.shape[0] <= k or k < 2 or k > 129):
        return params
    from irp_b200 import umap_graph
    knn_indices, knn_dists = umap_graph.nearest_neighbors(features_pca, k)
    params['precomputed_knn'] = (knn_indices, knn_dists)
    return params


def detect_outliers(embedding, labels, class_n_neighbors=30, class_contamination=0.05,
                    global_n_neighbors=75, global_contamination=0.03):
    """Detect outliers using Local Outlier Factor  (reference: data_curation.py:709-728)."""
    from sklearn.preprocessing import LabelEncoder

    le = LabelEncoder()
    y_numeric = le.fit_transform(labels)
    n_classes = len(le.classes_)
    device = torch.device(f"cuda:{torch.cuda.current_device()}")
    z = torch.from_numpy(np.ascontiguousarray(embedding, dtype=np.float32)).to(device)
    ids = torch.from_numpy(y_numeric.astype(np.int32)).to(device)
    counts = np.bincount(y_numeric, minlength=n_classes)
    if (counts <= class_n_neighbors).any() or len(labels) <= global_n_neighbors:
        # sklearn/neighbors/_lof.py:286-293 warns and clips k to n-1
        warnings.warn("n_neighbors is greater than the total number of samples in at least one group; "
                      "n_neighbors will be set to (n_samples - 1) for estimation.")
    _, _, cflags = ops.lof(z, ids, n_classes, class_n_neighbors, class_contamination)
    _, _, gflags = ops.lof(z, None, 1, global_n_neighbors, global_contamination)
    return cflags.cpu().numpy().astype(bool), gflags.cpu().numpy().astype(bool)


def create_results_dataframe(embedding, labels, paths, class_outliers, global_outliers):
    """Create comprehensive results dataframe  (reference: data_curation.py:730-743; host-side, unchanged)."""
    import pandas as pd
    from sklearn.preprocessing import LabelEncoder

    le = LabelEncoder()
    y_numeric = le.fit_transform(labels)
    return pd.DataFrame({
        'x': embedding[:, 0],
        'y': embedding[:, 1],
        'label': labels,
        'label_encoded': y_numeric,
        'path': paths,
        'is_class_outlier': class_outliers,
        'is_global_outlier': global_outliers
    })


# -----------------------------------------------------------------------------------------------------------------
# Dataset cleaning: the duplicate hash (reference functions/data_curation.py:283-292,
# call site :394-399)
# -----------------------------------------------------------------------------------------------------------------
def compute_image_hashes(images, device="cuda:0", batch_size=256):
    """Batch form of `compute_image_hash`: a list of PIL images -> list of md5 hex strings, identical to the
    reference's.  RGB images take the device path (Pillow-exact bicubic 64x64 resize + MD5, both libirp_b200 kernels).
    The reference resizes in the image's OWN mode and converts afterwards (:287-288); for the rare non-RGB inputs
    (the recorded dataset has 1 grayscale and 50 RGBA images in 26 179) that resize stays Pillow's on the host, like
    the decode, and the device computes the digest."""
    dev = torch.device(device)
    out = [None] * len(images)
    rgb = [i for i, im in enumerate(images) if im.mode == "RGB"]
    for lo in range(0, len(rgb), batch_size):
        idx = rgb[lo:lo + batch_size]
        part = pack_images([np.asarray(images[i]) for i in idx], transform=_lib.TRANSFORM_HASH_64).to(dev)
        digests = ops.image_hashes(part.pixels, part.offsets, part.hw, part.max_taps).cpu().numpy()
        for i, d in zip(idx, digests):
            out[i] = d.tobytes().hex()
    other = [i for i, im in enumerate(images) if im.mode != "RGB"]
    if other:
        small = np.stack([np.asarray(images[i].copy().resize((64, 64)).convert("RGB")) for i in other])
        data = torch.from_numpy(small.reshape(len(other), -1)).to(dev)
        for i, d in zip(other, ops.md5_rows(data).cpu().numpy()):
            out[i] = d.tobytes().hex()
    return out


def compute_image_hash(img, device="cuda:0"):
    """Compute a hash from image data to detect duplicates (drop-in: same hex digest as the reference)."""
    return compute_image_hashes([img], device=device)[0]


# -----------------------------------------------------------------------------------------------------------------
# WebDataset curation: the resize step (reference functions/data_curation.py:883-913)
# -----------------------------------------------------------------------------------------------------------------
def _to_rgb(img):
    """The reference's mode handling (data_curation.py:886-892): RGBA composited on white through its own alpha,
    any other non-RGB mode converted."""
    from PIL import Image

    if img.mode == 'RGBA':
        background = Image.new('RGB', img.size, (255, 255, 255))
        background.paste(img, mask=img.split()[3])
        return background
    if img.mode != 'RGB':
        return img.convert('RGB')
    return img


def resize_and_crop_images(images, device="cuda:0", batch_size=256):
    """Batch form of `resize_and_crop_image`: a list of PIL images -> uint8 array [n,224,224,3] (the pixels of the
    images the reference function returns), Lanczos resize + center crop on the device."""
    dev = torch.device(device)
    arrays = [np.asarray(_to_rgb(im)) for im in images]
    out = np.empty((len(arrays), _lib.CROP, _lib.CROP, 3), np.uint8)
    for lo in range(0, len(arrays), batch_size):
        part = pack_images(arrays[lo:lo + batch_size], transform=_lib.TRANSFORM_WDS_LANCZOS).to(dev)
        res = ops.preprocess_ex(part.pixels, part.offsets, part.hw, part.max_taps, _lib.LAYOUT_U8_HWC,
                                _lib.TRANSFORM_WDS_LANCZOS)
        out[lo:lo + len(res)] = res.cpu().numpy()
    return out


def resize_and_crop_image(img, target_size=224, device="cuda:0"):
    """Resize and center crop image to target_size x target_size (drop-in: returns a PIL image)."""
    from PIL import Image

    if target_size != _lib.CROP:
        raise ValueError("the B200 path is built for target_size=224 (the reference's only call site)")
    return Image.fromarray(resize_and_crop_images([img], device=device)[0])
