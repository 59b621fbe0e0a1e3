"""ctypes binding of include/sfm_b200.h. Fails loudly when the CUDA library is missing:
there is no CPU fallback, and the product never touches the test oracle."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# SFM_B200_LIB: development override used by tools/variants.py to time kernel build variants
LIB_PATH = os.environ.get("SFM_B200_LIB") or os.path.join(_HERE, "libsfm_b200.so")

SFM_OK = 0
SFM_E_INVALID = -1
SFM_E_NO_DEVICE = -2
SFM_E_CUDA = -3
SFM_E_DIM = -4
SFM_E_NOT_INTEGRAL = -5
SFM_E_RANGE = -6
SFM_E_TOO_FEW_TRAIN = -7
SFM_E_CAPACITY = -8
SFM_E_NOT_UPLOADED = -9
SFM_E_NOMEM = -10

# every symbol include/sfm_b200.h declares: (restype, argtypes)
_vp, _i, _i64, _f, _d = C.c_void_p, C.c_int, C.c_int64, C.c_float, C.c_double
_pi, _pi64, _pf, _pd = C.POINTER(C.c_int32), C.POINTER(C.c_int64), C.POINTER(C.c_float), C.POINTER(C.c_double)
SYMBOLS = {
    "sfm_abi_version": (_i, []),
    "sfm_create": (_vp, [_i, C.POINTER(C.c_int)]),
    "sfm_destroy": (None, [_vp]),
    "sfm_last_error": (C.c_char_p, [_vp]),
    "sfm_strerror": (C.c_char_p, [_i]),
    "sfm_host_alloc": (_vp, [C.c_size_t]),
    "sfm_host_free": (None, [_vp]),
    "sfm_upload_descriptors": (_i, [_vp, _i, C.POINTER(_vp), _pi, _i]),
    "sfm_upload_descriptors_u8": (_i, [_vp, _i, C.POINTER(_vp), _pi, _i]),
    "sfm_upload_descriptors_async": (_i, [_vp, _i, C.POINTER(_vp), _pi, _i, _i]),
    "sfm_upload_descriptors_bin": (_i, [_vp, _i, C.POINTER(_vp), _pi, _i]),
    "sfm_reproject_jacobians": (_i, [_vp, _pd, _pd, _i, _pd, _i64, _pi, _pi, _pf, _i64, _pd, _pd, _i, _pf]),
    "sfm_probe_fp64_peak": (_i, [_vp, _i, C.POINTER(C.c_double)]),
    "sfm_estimate_normals": (_i, [_vp, _pd, _i64, _i, _pd]),
    "sfm_upload_keypoints": (_i, [_vp, _i, C.POINTER(_vp), _pi]),
    "sfm_get_matched_points": (_i, [_vp, _i, C.POINTER(C.c_uint8), _pf, _pf, _i64, _pi64]),
    "sfm_reconstruct_pair": (_i, [_vp, _i, _pd, _pd, _pd, _pd, _pd, C.POINTER(C.c_uint8), _pd, _i64, _pi64]),
    "sfm_save_structure": (_i, [C.c_char_p, _i, _pd, _pd, _i64, _pd, _i64, C.POINTER(C.c_uint8)]),
    "sfm_write_ply_binary": (_i, [C.c_char_p, _i64, _pf, C.POINTER(C.c_uint8), _i]),
    "sfm_match_pairs": (_i, [_vp, _pi, _pi, _i, _d, _f, _f, _vp, _i64, _pi64, _vp, _pf]),
    "sfm_fetch_matches": (_i, [_vp, _vp, _i64]),
    "sfm_match_pairs_resident": (_i, [_vp, _pi, _pi, _i, _d, _f, _f, _pi64, _pf, _pf]),
    "sfm_triangulate_batch": (_i, [_vp, _pf, _pf, _i, _i64, _pf, _pd]),
    "sfm_reproject_residuals": (_i, [_vp, _pd, _pd, _i, _pd, _i64, _pi, _pi, _pf, _i64, _d, _pd, _pd]),
    "sfm_triangulate_batch_timed": (_i, [_vp, _pf, _pf, _i, _i64, _pf, _pd, _i, _pf]),
    "sfm_reproject_residuals_timed": (_i, [_vp, _pd, _pd, _i, _pd, _i64, _pi, _pi, _pf, _i64, _d, _pd, _pd, _i, _pf]),
    "sfm_probe_i8_peak": (_i, [_vp, _i, _pd]),
    "sfm_launch_count": (_i64, [_vp]),
    "sfm_last_rechecked_rows": (_i64, [_vp]),
    "sfm_timer_start": (_i, [_vp]),
    "sfm_timer_stop": (_i, [_vp, _pf]),
    "sfm_sync": (_i, [_vp]),
    "sfm_bank_layout": (_i, [_vp, _i, _pi, _i]),
    "sfm_bank_upload_range": (_i, [_vp, _i, _i, C.POINTER(_vp), _i]),
    "sfm_bank_commit": (_i, [_vp, _i, _i]),
    "sfm_bank_image_rows": (_i, [_vp, _i, _pi64, _pi64]),
    "sfm_bank_rows_dev": (_vp, [_vp, _pi64]),
    "sfm_bank_copy_peer": (_i, [_vp, _vp, _i, _i]),
    "sfm_bank_layout_async": (_i, [_vp, _i, _pi, _i]),
    "sfm_bank_upload_range_async": (_i, [_vp, _i, _i, C.POINTER(_vp), _i]),
    "sfm_bank_commit_async": (_i, [_vp, _i, _i]),
    "sfm_upload_stream": (_vp, [_vp]),
    "sfm_peer_export": (_i, [_vp, _vp]),
    "sfm_peer_connect": (_i, [_vp, _i, _i, _vp]),
    "sfm_peer_disconnect": (_i, [_vp]),
    "sfm_bank_ready_async": (_i, [_vp, C.c_uint32]),
    "sfm_bank_push_range_async": (_i, [_vp, _i, _i, _i, C.c_uint32]),
    "sfm_bank_pull_commit_async": (_i, [_vp, _i, _i, _i, _i, C.c_uint32]),
    "sfm_match_rows_begin": (_i, [_vp, _pi, _pi, _pi, _pi, _i, _d, _pf]),
    "sfm_match_rows_finish": (_i, [_vp, _pf, _f, _f, _pi64, _vp]),
    "sfm_ba_create": (_i, [_vp, _i, _i64, _pi, _pi, _pf, _i64, C.POINTER(_vp)]),
    "sfm_ba_evaluate": (_i, [_vp, _vp, _pd, _pd, _pd, _d, _pd, _pd, _pd, _pf]),
    "sfm_ba_destroy": (None, [_vp, _vp]),
}

_lib = None


def load() -> C.CDLL:
    """Loads libsfm_b200.so (built in-tree by sfm_opencv_b200/build.py)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -m sfm_opencv_b200.build` "
            "(nvcc, sm_100a). There is no CPU or PyTorch fallback for this path.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)          # AttributeError if the ABI lost a symbol
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


class SfmError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"sfm_b200 error {code}: {msg}")
        self.code = code
