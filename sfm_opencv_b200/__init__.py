"""sfm_opencv_b200 -- B200-native (sm_100a) hot path of CaptainEven/SFM_OpenCV:
pairwise SIFT kNN-2 + ratio matching, batched DLT triangulation, reprojection residuals.
The compute lives in libsfm_b200.so (hand-written CUDA behind include/sfm_b200.h)."""
from ._capi import LIB_PATH, SYMBOLS, SfmError, load  # noqa: F401
from .api import (BAProblem, Context, KNN_DTYPE, MATCH_DTYPE, build_projection,  # noqa: F401
                  bundle_adjustment_residuals, enumerate_observations, match_features,
                  match_features_for_all, reconstruct, save_structure, write_ply_binary)

__all__ = ["Context", "BAProblem", "SfmError", "match_features", "match_features_for_all", "reconstruct",
           "build_projection", "enumerate_observations", "bundle_adjustment_residuals", "save_structure", "write_ply_binary", "load", "LIB_PATH", "SYMBOLS", "MATCH_DTYPE", "KNN_DTYPE"]
