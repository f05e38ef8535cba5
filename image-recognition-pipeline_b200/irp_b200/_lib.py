"""ctypes binding of libirp_b200.so (the C ABI declared in include/irp_b200.h).

The library is built in-tree (image-recognition-pipeline_b200/lib/libirp_b200.so) by
``__graft_entry__.build()`` / ``make -C image-recognition-pipeline_b200/csrc``.  There is no Python or CPU
fallback: if the shared object is missing, or the device is not sm_100, loading fails loudly.
"""
from __future__ import annotations

import ctypes as C
import os
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.normpath(os.path.join(_HERE, "..", "lib", "libirp_b200.so"))

IRP_OK = 0
PCA_SOLVER_AUTO, PCA_SOLVER_LANCZOS, PCA_SOLVER_HOUSEHOLDER = 0, 1, 2
LAYOUT_NCHW = 0
TRANSFORM_WEIGHTS_DEFAULT = 0  # ResNet50_Weights.DEFAULT.transforms() (resize 232, crop 224)
TRANSFORM_VAL_256 = 1          # functions/dataload.py:51-56 (Resize((256,256)), CenterCrop(224))
TRANSFORM_WDS_LANCZOS = 2      # functions/data_curation.py:883-913 (smaller side -> 224, LANCZOS, center crop)
TRANSFORM_HASH_64 = 3          # functions/data_curation.py:283-292 (img.resize((64, 64)), BICUBIC) -> uint8 [n,64,64,3]
HASH_SIZE = 64
LAYOUT_U8_HWC = 2              # uint8 [n,224,224,3]: the resized pixels themselves (no normalisation)
LAYOUT_NHWC4P = 1
CROP = 224
PAD_HW = 230
EMBED_DIM = 2048
NUM_CONVS = 53

_vp = C.c_void_p
_i = C.c_int
_i64 = C.c_int64
_sz = C.c_size_t
_f = C.c_float
_d = C.c_double

# name -> (restype, argtypes); every symbol include/irp_b200.h declares
SIGNATURES = {
    "irp_abi_version": (_i, []),
    "irp_last_error": (C.c_char_p, []),
    "irp_build_id": (C.c_char_p, []),
    "irp_init": (_i, [_i]),
    "irp_preprocess_geometry": (_i, [_i, _i] + [C.POINTER(_i)] * 5),
    "irp_preprocess_workspace_bytes": (_sz, [_i, _i]),
    "irp_preprocess": (_i, [_vp, _vp, _vp, _i, _i, _vp, _sz, _vp, _i, _vp]),
    "irp_preprocess_status": (_i, [_vp, _i, _i, _vp]),
    "irp_preprocess_ex": (_i, [_vp, _vp, _vp, _i, _i, _vp, _sz, _vp, _i, _i, _vp]),
    "irp_preprocess_geometry_ex": (_i, [_i, _i, _i] + [C.POINTER(_i)] * 5),
    "irp_md5_rows": (_i, [_vp, _i, _i64, _vp, _vp]),
    "irp_classifier_head_workspace_bytes": (_sz, [_i, _i]),
    "irp_classifier_head": (_i, [_vp, _i, _i, _vp, _vp, _i, _vp, _vp, _i, _vp, _vp, _vp, _sz, _vp]),
    "irp_cross_entropy_stats": (_i, [_vp, _vp, _i, _i, _vp, _vp, _vp]),
    "irp_resnet50_create": (_i, [C.POINTER(_vp), _i]),
    "irp_resnet50_destroy": (None, [_vp]),
    "irp_resnet50_conv_shape": (_i, [_i] + [C.POINTER(_i)] * 5),
    "irp_resnet50_load_conv": (_i, [_vp, _i, _vp, _vp, _vp, _vp, _vp, _f, _vp]),
    "irp_resnet50_embed": (_i, [_vp, _vp, _i, _vp, _vp]),
    "irp_resnet50_embed_capture": (_i, [_vp, _vp, _i, _vp, _i, _vp, _sz, _vp]),
    "irp_conv2d_nhwc": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _i, _vp]),
    "irp_conv1x1_chain": (_i, [_vp] * 8 + [_i64, _i, _i, _i, _vp]),
    "irp_conv1x1_pool": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _vp]),
    "irp_conv1x1_chain_ds": (_i, [_vp] * 8 + [_i64, _i, _i, _i, _i, _vp]),
    "irp_cov_workspace_bytes": (_sz, [_i64, _i]),
    "irp_cov_accumulate": (_i, [_vp, _i64, _i, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "irp_pca_fit_workspace_bytes": (_sz, [_i, _i]),
    "irp_pca_fit": (_i, [_vp, _vp, _vp, _vp, _i, _i, _vp, _vp, _vp, _vp, _sz, _vp]),
    "irp_pca_fit_ex": (_i, [_vp, _vp, _vp, _vp, _i, _i, _vp, _vp, _vp, _vp, _sz, _i, _i, C.POINTER(C.c_int32), _vp]),
    "irp_pca_transform_workspace_bytes": (_sz, [_i64, _i, _i]),
    "irp_pca_transform": (_i, [_vp, _i64, _i, _vp, _vp, _i, _vp, _vp, _sz, _vp]),
    "irp_lof_workspace_bytes": (_sz, [_i64, _i, _i]),
    "irp_lof": (_i, [_vp, _i64, _i, _vp, _i, _i, _d, _vp, _vp, _vp, _vp, _sz, _vp]),
    "irp_lof_knn_part": (_i, [_vp, _i64, _i, _vp, _i, _i, _i, _i, _vp, _vp, _sz, _vp]),
    "irp_lof_lrd_part": (_i, [_i64, _i, _i, _i, _i, _vp, _vp, _vp, _sz, _vp]),
    "irp_lof_score_part": (_i, [_i64, _i, _i, _i, _i, _vp, _vp, _vp, _sz, _vp]),
    "irp_lof_finish": (_i, [_i64, _i, _i, _d, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "irp_knn_graph_workspace_bytes": (_sz, [_i64, _i, _i]),
    "irp_knn_graph": (_i, [_vp, _i64, _i, _i, _vp, _vp, _vp, _sz, _vp]),
    "irp_umap_fuzzy_weights": (_i, [_vp, _vp, _i64, _i, _f, _f, _i, _vp, _vp, _vp, _vp, _sz, _vp]),
    "irp_centroid_workspace_bytes": (_sz, [_i64, _i, _i]),
    "irp_centroid_zscore": (_i, [_vp, _i64, _i, _vp, _i, _d, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
}

_lock = threading.Lock()
_lib = None
_inited_devices = set()


class IrpError(RuntimeError):
    """A libirp_b200 call returned a non-zero status (message from irp_last_error())."""


def load() -> C.CDLL:
    """dlopen the in-tree shared object and attach the prototypes. Raises if it has not been built."""
    global _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(there is no CPU fallback for the outlier-stage hot path)"
            )
        lib = C.CDLL(LIB_PATH)
        partial = os.environ.get("IRP_B200_PARTIAL") == "1"  # developer probes against a half-built library
        for name, (res, args) in SIGNATURES.items():
            if partial and not hasattr(lib, name):
                continue
            fn = getattr(lib, name)  # AttributeError if the .so is stale
            fn.restype = res
            fn.argtypes = args
        _lib = lib
        return lib


def source_build_id() -> str:
    """The id `make` would stamp into the library for the sources currently on disk (see csrc/Makefile)."""
    import glob
    import hashlib
    csrc = os.path.normpath(os.path.join(_HERE, "..", "csrc"))
    files = sorted(glob.glob(os.path.join(csrc, "*.cu")) + glob.glob(os.path.join(csrc, "*.cuh")) +
                   glob.glob(os.path.join(csrc, "*.h")), key=os.path.basename)
    files.append(os.path.normpath(os.path.join(_HERE, "..", "..", "include", "irp_b200.h")))
    h = hashlib.sha256()
    for f in files:
        with open(f, "rb") as fh:
            h.update(fh.read())
    return h.hexdigest()[:16]


def assert_fresh() -> str:
    """Raise if the loaded library was not built from the sources beside it; returns the build id."""
    built = load().irp_build_id().decode()
    want = source_build_id()
    if built != want:
        raise ImportError(f"{LIB_PATH} is stale: built from sources {built}, sources on disk are {want}; rebuild with "
                          "`python -c 'import __graft_entry__ as g; g.build()'`")
    return built


def check(status: int, what: str = "") -> None:
    if status != IRP_OK:
        msg = load().irp_last_error()
        raise IrpError(f"{what or 'libirp_b200'} failed with status {status}: {msg.decode() if msg else ''}")


def init(device_index: int) -> C.CDLL:
    """Load the library and validate `device_index` (must be a compute-capability 10.x GPU)."""
    lib = load()
    if device_index not in _inited_devices:
        check(lib.irp_init(int(device_index)), "irp_init")
        _inited_devices.add(device_index)
    return lib


def geometry(h: int, w: int, transform: int = 0):
    """(out_h, out_w, top, left, taps) of the resize / crop-224 transform for an h x w image
    (transform 0: ResNet50_Weights.DEFAULT.transforms(), 1: the classifier's Resize((256,256)) val_transform)."""
    lib = load()
    vals = [C.c_int() for _ in range(5)]
    check(lib.irp_preprocess_geometry_ex(int(h), int(w), int(transform), *[C.byref(v) for v in vals]),
          "irp_preprocess_geometry_ex")
    return tuple(v.value for v in vals)
