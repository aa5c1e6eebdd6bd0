"""ctypes binding of libgsf.so (include/gsf.h).  Fails loudly when the library is missing:
there is no CPU fallback anywhere in this package."""
from __future__ import annotations

import ctypes
import os
from ctypes import c_char_p, c_double, c_int32, c_int64, c_uint64, c_void_p

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("GSF_LIB") or os.path.join(PKG_DIR, "libgsf.so")     # GSF_LIB: tuning hook (A/B builds)

GSF_E_INVALID, GSF_E_CUDA, GSF_E_NO_DEVICE, GSF_E_TOO_LARGE = -1, -2, -3, -4
ST_OK, ST_TOO_FEW_POINTS, ST_DEGENERATE, ST_BAD_QUATERNION = 0, 1, 2, 4
ST_EMPTY, ST_RANSAC_OUTLIERS, ST_TOO_LONG, ST_GRID_NEEDS_ALL_VALID, ST_NEEDS_FP64 = 8, 16, 32, 64, 128
GEO_PARTS = 1024

# name -> (restype, argtypes); the CPU test-suite checks that every symbol declared in
# include/gsf.h is listed here and exported by the library.
SIGNATURES = {
    "gsf_version": (c_char_p, []),
    "gsf_last_error": (c_char_p, []),
    "gsf_device_sm_count": (c_int32, []),
    "gsf_fuse_batched_dev": (c_int32, [c_void_p] * 5 + [c_int32, c_int64, c_void_p, c_int32] + [c_void_p] * 7),
    "gsf_ekf_strict_batched_dev": (c_int32, [c_void_p] * 5 + [c_int32, c_void_p, c_int32] + [c_void_p] * 6),
    "gsf_hypothesis_grid_work_doubles": (c_int64, [c_int64, c_int32]),
    "gsf_ekf_hypothesis_grid_dev": (c_int32, [c_void_p] * 4 + [c_int64, c_void_p, c_int32] + [c_void_p] * 5),
    "gsf_noise_grid_work_doubles": (c_int64, [c_int64, c_int32, c_int32, c_int32, c_int64, c_int64]),
    "gsf_ekf_noise_grid_dev": (c_int32, [c_void_p] * 4 + [c_int64] + [c_void_p] * 4 + [c_int32] * 3 + [c_int64] * 2 + [c_void_p] * 5),
    "gsf_poly_ransac_dev": (c_int32, [c_void_p, c_void_p, c_int32] + [c_void_p] * 6 + [c_int32] * 4 + [c_double] + [c_void_p] * 4),
    "gsf_parse_table_work_bytes": (c_int64, [c_int64]),
    "gsf_parse_table_dev": (c_int32, [c_void_p, c_int64, c_int32, c_int32, c_void_p, c_int64, c_void_p, c_void_p, c_void_p]),
    "gsf_write_rows_work_bytes": (c_int64, [c_int64]),
    "gsf_write_pose_rows_dev": (c_int32, [c_void_p, c_void_p, c_void_p, c_int64, c_void_p, c_char_p, c_int32, c_void_p, c_int64,
                                          c_void_p, c_void_p, c_void_p]),
    "gsf_to_local_f32_dev": (c_int32, [c_void_p] * 6 + [c_int32] + [c_void_p] * 6),
    "gsf_from_local_f32_dev": (c_int32, [c_void_p] * 4 + [c_int32] + [c_void_p] * 4),
    "gsf_fuse_batched_f32_dev": (c_int32, [c_void_p] * 7 + [c_int32, c_void_p, c_int32] + [c_void_p] * 5),
    "gsf_associate_spline_long_work_doubles": (c_int64, [c_int64, c_int64]),
    "gsf_associate_spline_long_dev": (c_int32, [c_void_p, c_void_p, c_int64, c_void_p, c_int64, c_double] + [c_void_p] * 5),
    "gsf_ekf_step_dev": (c_int32, [c_int32] + [c_void_p] * 9 + [c_int32] + [c_void_p] * 6),
    "gsf_rts_segment_dev": (c_int32, [c_void_p] * 5 + [c_int32] + [c_void_p] * 3),
    "gsf_quat_nlerp_dev": (c_int32, [c_void_p] * 3 + [c_int64, c_void_p, c_void_p]),
    "gsf_sharp_turn_dev": (c_int32, [c_void_p] * 3 + [c_int32, c_double] + [c_void_p] * 3),
    "gsf_umeyama_work_doubles": (c_int64, [c_int32, c_int64]),
    "gsf_sim3_partial_stats_dev": (c_int32, [c_void_p] * 4 + [c_int64] + [c_void_p] * 3),
    "gsf_sim3_from_partial_stats_dev": (c_int32, [c_void_p, c_int32] + [c_void_p] * 5),
    "gsf_sim3_umeyama_batched_dev": (c_int32, [c_void_p] * 4 + [c_int32, c_int64] + [c_void_p] * 6),
    "gsf_sim3_ransac_work_doubles": (c_int64, [c_int32, c_int64]),
    "gsf_sim3_ransac_dev": (c_int32, [c_void_p, c_void_p, c_int64, c_void_p, c_int32, c_int32, c_double, c_int32] + [c_void_p] * 8),
    "gsf_sim3_apply_dev": (c_int32, [c_void_p] * 6 + [c_int32, c_int64] + [c_void_p] * 4),
    "gsf_ate_work_doubles": (c_int64, [c_int64, c_int64]),
    "gsf_ate_nn_batched_dev": (c_int32, [c_void_p] * 4 + [c_int32, c_int64, c_double, c_void_p, c_void_p, c_void_p]),
    "gsf_utm_forward_dev": (c_int32, [c_void_p, c_void_p, c_int64, c_int32, c_int32, c_void_p, c_void_p, c_void_p]),
    "gsf_utm_inverse_dev": (c_int32, [c_void_p, c_void_p, c_int64, c_int32, c_int32, c_void_p, c_void_p, c_void_p]),
    "gsf_gnss_rows_to_utm_dev": (c_int32, [c_void_p, c_int64] + [c_void_p] * 5),
    "gsf_geo_zone_dev": (c_int32, [c_void_p, c_void_p, c_int64, c_void_p, c_void_p, c_void_p]),
    "gsf_associate_spline_dev": (c_int32, [c_void_p] * 5 + [c_int32, c_double] + [c_void_p] * 4),
    "gsf_synth_generate_dev": (c_int32, [c_void_p] * 4 + [c_int64, c_int32, c_int32, c_double, c_double, c_uint64,
                                         c_double, c_int32, c_void_p]),
    "gsf_fuse_batched_host": (c_int32, [c_void_p] * 5 + [c_int32, c_int64, c_void_p, c_int32] + [c_void_p] * 6),
    "gsf_host_workspace_free": (None, []),
}

_lib = None


class GsfError(RuntimeError):
    pass


def load(build_if_missing: bool = False) -> ctypes.CDLL:
    """Load libgsf.so.  Raises if it has not been built (``python -m gps_optimize_slam_b200.build``)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        if build_if_missing:
            from . import build as _build
            _build.build()
        else:
            raise GsfError(
                f"{LIB_PATH} is missing: the CUDA library has not been built "
                "(run `python -m gps_optimize_slam_b200.build`).  There is no CPU fallback.")
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = load().gsf_last_error().decode("utf-8", "replace")
        raise GsfError(f"{what} failed (code {rc}): {msg}")
