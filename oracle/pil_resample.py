"""ORACLE (test infrastructure, never on the product path): numpy restatement of the reference's image transform.

The reference applies ``ResNet50_Weights.DEFAULT.transforms()`` to a PIL image
(/root/reference/functions/data_curation.py:675).  That preset is
torchvision/transforms/_presets.py ``ImageClassification.forward`` (crop 224, resize 232, bilinear,
antialias) = resize -> center_crop -> pil_to_tensor -> /255 -> normalize, with

* output size: torchvision/transforms/functional.py:359-384 (`_compute_resized_output_size`): short side -> 232,
  long side -> int(232 * long / short);
* the resize itself: Pillow ``Image.resize(..., BILINEAR)`` = ``ImagingResample`` for 8-bit images
  (Pillow src/libImaging/Resample.c, not vendored under /root/reference; pinned Pillow 11.0.0, container 12.2.0):
  ``precompute_coeffs`` (triangle filter, support = max(scale, 1)), ``normalize_coeffs_8bpc``
  (PRECISION_BITS = 32 - 8 - 2 = 22), horizontal pass then vertical pass, each ``clip8((acc + 2^21) >> 22)``;
* crop offsets: functional.py:592-593 ``int(round((H' - 224) / 2.0))`` (Python round-half-even).

Pinned: tests/test_oracle.py checks this module bit-for-bit against Pillow/torchvision run in-process and
against tests/golden/preprocess.npz (generated from the reference transform by oracle/make_golden.py).
"""
from __future__ import annotations

import math

import numpy as np

PRECISION_BITS = 22
RESIZE = 232
CROP = 224
MEAN = np.array([0.485, 0.456, 0.406], dtype=np.float32)
STD = np.array([0.229, 0.224, 0.225], dtype=np.float32)


def resized_size(h: int, w: int) -> tuple[int, int]:
    """(out_h, out_w) of F.resize(img, [232]) -- functional.py:359-384."""
    if w <= h:
        return int(RESIZE * h / w), RESIZE
    return RESIZE, int(RESIZE * w / h)


def crop_offsets(out_h: int, out_w: int) -> tuple[int, int]:
    """(top, left) of center_crop(224) -- functional.py:592-593 (Python round = half to even)."""
    return int(round((out_h - CROP) / 2.0)), int(round((out_w - CROP) / 2.0))


def max_taps(h: int, w: int) -> int:
    out_h, out_w = resized_size(h, w)
    s = max(h / out_h, w / out_w, 1.0)
    return 2 * int(math.ceil(s)) + 1


def _triangle(x: float) -> float:
    x = abs(x)
    return 1.0 - x if x < 1.0 else 0.0


def _sinc(x: float) -> float:
    if x == 0.0:
        return 1.0
    x = x * math.pi
    return math.sin(x) / x


def _lanczos3(x: float) -> float:
    """Pillow lanczos_filter: truncated sinc, support 3."""
    if -3.0 <= x < 3.0:
        return _sinc(x) * _sinc(x / 3.0)
    return 0.0


def _bicubic(x: float) -> float:
    """Pillow bicubic_filter (a = -0.5), support 2."""
    a = -0.5
    x = abs(x)
    if x < 1.0:
        return ((a + 2.0) * x - (a + 3.0)) * x * x + 1
    if x < 2.0:
        return (((x - 5) * x + 8) * x - 4) * a
    return 0.0


FILTERS = {"bilinear": (_triangle, 1.0), "lanczos": (_lanczos3, 3.0), "bicubic": (_bicubic, 2.0)}


def precompute_coeffs(in_size: int, out_size: int, filt: str = "bilinear"):
    """Pillow precompute_coeffs + normalize_coeffs_8bpc for the bilinear (triangle) or the Lanczos filter.

    Returns (first[out_size] int, count[out_size] int, coef[out_size, ksize] int64 fixed point)."""
    weight, base_support = FILTERS[filt]
    scale = in_size / out_size
    filterscale = max(scale, 1.0)
    support = base_support * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    first = np.zeros(out_size, np.int64)
    count = np.zeros(out_size, np.int64)
    coef = np.zeros((out_size, ksize), np.int64)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = (xx + 0.5) * scale
        xmin = max(int(center - support + 0.5), 0)
        xmax = min(int(center + support + 0.5), in_size)
        n = xmax - xmin
        ws = []
        ww = 0.0
        for x in range(n):
            wgt = weight((x + xmin - center + 0.5) * ss)
            ws.append(wgt)
            ww += wgt
        for x in range(n):
            wgt = ws[x] / ww if ww != 0.0 else ws[x]
            coef[xx, x] = int(-0.5 + wgt * (1 << PRECISION_BITS)) if wgt < 0 else int(0.5 + wgt * (1 << PRECISION_BITS))
        first[xx] = xmin
        count[xx] = n
    return first, count, coef


def _clip8(acc: np.ndarray) -> np.ndarray:
    return np.clip(acc >> PRECISION_BITS, 0, 255).astype(np.uint8)


def _pass(img: np.ndarray, axis: int, out_size: int, filt: str = "bilinear") -> np.ndarray:
    """One separable pass along `axis` (0 = rows/vertical, 1 = cols/horizontal) rounded to uint8 like Pillow."""
    in_size = img.shape[axis]
    first, count, coef = precompute_coeffs(in_size, out_size, filt)
    src = np.moveaxis(img, axis, 0).astype(np.int64)
    out = np.empty((out_size,) + src.shape[1:], np.uint8)
    for o in range(out_size):
        n = int(count[o])
        f = int(first[o])
        acc = np.tensordot(coef[o, :n], src[f:f + n], axes=(0, 0)) + (1 << (PRECISION_BITS - 1))
        out[o] = _clip8(acc)
    return np.moveaxis(out, 0, axis)


def resize_bilinear_u8(img: np.ndarray, out_h: int, out_w: int, filt: str = "bilinear") -> np.ndarray:
    """Pillow Image.resize((out_w, out_h), BILINEAR | LANCZOS) on an HWC uint8 array (horizontal pass first)."""
    h, w = img.shape[:2]
    out = img
    if w != out_w:
        out = _pass(out, 1, out_w, filt)
    if h != out_h:
        out = _pass(out, 0, out_h, filt)
    return out


def transform_u8(img: np.ndarray) -> np.ndarray:
    """resize(232) + center_crop(224) -> uint8 [224,224,3] (the pixels entering to_tensor/normalize)."""
    h, w = img.shape[:2]
    out_h, out_w = resized_size(h, w)
    r = img if (out_h, out_w) == (h, w) else resize_bilinear_u8(img, out_h, out_w)
    top, left = crop_offsets(out_h, out_w)
    return np.ascontiguousarray(r[top:top + CROP, left:left + CROP])


def normalize(u8_hwc: np.ndarray) -> np.ndarray:
    """pil_to_tensor -> float32 / 255 -> (x - mean) / std, CHW float32 (exact float32 arithmetic of torch)."""
    x = u8_hwc.astype(np.float32) / np.float32(255.0)
    x = (x - MEAN) / STD
    return np.ascontiguousarray(x.transpose(2, 0, 1))


def transform(img: np.ndarray) -> np.ndarray:
    """The whole reference transform on an HWC uint8 RGB array -> float32 [3,224,224]."""
    return normalize(transform_u8(img))


# ---------------------------------------------------------------------------------------------------------------
# the classifier's validation transform: /root/reference/functions/dataload.py:51-56
#   Resize((256, 256)) -> CenterCrop(224) -> ToTensor -> Normalize(ImageNet mean/std)
# Resize with a (h, w) pair ignores the aspect ratio (torchvision/transforms/functional.py:476-478 -> PIL
# Image.resize, same antialiased bilinear ImagingResample as above); ToTensor = uint8 -> float32 / 255
# (functional.py:169-175), Normalize = (x - mean) / std (functional.py `normalize`).
# ---------------------------------------------------------------------------------------------------------------
VAL_RESIZE = 256


def val_max_taps(h: int, w: int) -> int:
    s = max(h / VAL_RESIZE, w / VAL_RESIZE, 1.0)
    return 2 * int(math.ceil(s)) + 1


def val_transform_u8(img: np.ndarray) -> np.ndarray:
    """Resize((256,256)) + CenterCrop(224) -> uint8 [224,224,3]."""
    h, w = img.shape[:2]
    r = img if (h, w) == (VAL_RESIZE, VAL_RESIZE) else resize_bilinear_u8(img, VAL_RESIZE, VAL_RESIZE)
    top, left = crop_offsets(VAL_RESIZE, VAL_RESIZE)
    return np.ascontiguousarray(r[top:top + CROP, left:left + CROP])


def val_transform(img: np.ndarray) -> np.ndarray:
    """The reference val_transform on an HWC uint8 RGB array -> float32 [3,224,224]."""
    return normalize(val_transform_u8(img))


# ---------------------------------------------------------------------------------------------------------------
# the WebDataset stage's resize: /root/reference/functions/data_curation.py:883-913
# resize_and_crop_image: RGBA is composited on white / other modes converted to RGB (host side, not restated here),
# smaller side -> 224 and the other one int(side * (224 / smaller)), Image.resize(..., LANCZOS), crop by floor
# division, returns the uint8 image.
# ---------------------------------------------------------------------------------------------------------------
def wds_resized_size(h: int, w: int, target: int = CROP) -> tuple[int, int]:
    if w < h:
        return int(h * (target / w)), target
    return target, int(w * (target / h))


def wds_max_taps(h: int, w: int) -> int:
    out_h, out_w = wds_resized_size(h, w)
    s = 3.0 * max(h / out_h, w / out_w, 1.0)
    return 2 * int(math.ceil(s)) + 1


def wds_transform_u8(img: np.ndarray) -> np.ndarray:
    """resize_and_crop_image on an HWC uint8 RGB array -> uint8 [224,224,3]."""
    h, w = img.shape[:2]
    out_h, out_w = wds_resized_size(h, w)
    r = resize_bilinear_u8(img, out_h, out_w, "lanczos")
    top, left = (out_h - CROP) // 2, (out_w - CROP) // 2
    return np.ascontiguousarray(r[top:top + CROP, left:left + CROP])


# ---------------------------------------------------------------------------------------------------------------
# the duplicate-detection hash: /root/reference/functions/data_curation.py:283-292
# compute_image_hash: img.resize((64, 64)) -- Pillow's default filter for RGB images is BICUBIC, aspect ratio not
# kept, no crop -- then convert("RGB") (a no-op for RGB input), tobytes(), hashlib.md5(...).hexdigest().
# ---------------------------------------------------------------------------------------------------------------
HASH_SIZE = 64


def hash_max_taps(h: int, w: int) -> int:
    s = 2.0 * max(h / HASH_SIZE, w / HASH_SIZE, 1.0)
    return 2 * int(math.ceil(s)) + 1


def hash_resize_u8(img: np.ndarray) -> np.ndarray:
    """Image.resize((64, 64)) of an HWC uint8 RGB array -> uint8 [64,64,3]."""
    return np.ascontiguousarray(resize_bilinear_u8(img, HASH_SIZE, HASH_SIZE, "bicubic"))


def image_hash(img: np.ndarray) -> str:
    """compute_image_hash of an RGB image given as an HWC uint8 array."""
    import hashlib
    return hashlib.md5(hash_resize_u8(img).tobytes()).hexdigest()


def to_bf16_bits(x: np.ndarray) -> np.ndarray:
    """float32 -> bfloat16 (round to nearest even), returned as uint16 bit patterns."""
    u = np.ascontiguousarray(x, np.float32).view(np.uint32).astype(np.uint64)
    rounded = (u + 0x7FFF + ((u >> 16) & 1)) >> 16
    return rounded.astype(np.uint16)
