"""CONFIG of the fusion path -- same keys and values as the reference's module-level dict
(EKFGPSSLAM.py:22-71) -- and its packing into the C-ABI ``FuseParams`` record
(include/gsf.h, gps_optimize_slam_b200/csrc/gsf_common.cuh)."""
from __future__ import annotations

import copy
import struct

import numpy as np

CONFIG = {
    "ekf": {
        "initial_cov_diag": [0.1, 0.1, 0.1, 0.01, 0.01, 0.01, 0.01],
        "process_noise_diag": [0.1, 0.1, 0.7, 0.01, 0.01, 0.01, 0.01],
        "meas_noise_diag": [0.2, 0.2, 0.2],
        "transition_steps": 10,
    },
    "sim3_ransac": {
        "min_samples": 4,
        "residual_threshold": 4.0,
        "max_trials": 1000,
        "min_inliers_needed": 4,
        "max_initial_duration": 180.0,
    },
    "gps_filtering_ransac": {
        "enabled": True,
        "use_sliding_window": True,
        "window_duration_seconds": 15.0,
        "window_step_factor": 0.5,
        "polynomial_degree": 2,
        "min_samples": 6,
        "residual_threshold_meters": 10.0,
        "max_trials": 50,
    },
    "time_alignment": {
        "max_samples_for_corr": 500,
        "max_gps_gap_threshold": 5.0,
    },
    "ground_truth_gps_filtering": {
        "enabled": False,
        "use_sliding_window": True,
        "window_duration_seconds": 15.0,
        "window_step_factor": 0.5,
        "polynomial_degree": 2,
        "min_samples": 6,
        "residual_threshold_meters": 5.0,
        "max_trials": 50,
    },
    "rts_decision": {
        "sharp_turn_yaw_rate_threshold_deg_per_sec": 45.0,
        "default_ekf_transition_steps_on_sharp_turn": 0,
    },
}

FUSE_PARAMS_BYTES = 23 * 8
EVAL_SKIP_SECONDS = 5.0          # EKFGPSSLAM.py:1021


def default_config():
    return copy.deepcopy(CONFIG)


def pack_fuse_params(cfg=None, *, p0=None, q=None, r=None) -> np.ndarray:
    """-> uint8[184] holding one ``FuseParams``.  ``p0``/``q``/``r`` override the EKF noise
    diagonals (per-trajectory / per-hypothesis parameter sets)."""
    cfg = CONFIG if cfg is None else cfg
    ekf, rc, ta, rts = cfg["ekf"], cfg["sim3_ransac"], cfg["time_alignment"], cfg["rts_decision"]
    p0 = list(ekf["initial_cov_diag"] if p0 is None else p0)
    q = list(ekf["process_noise_diag"] if q is None else q)
    r = list(ekf["meas_noise_diag"] if r is None else r)
    if len(p0) != 7 or len(q) != 7 or len(r) != 3:
        raise ValueError("EKF noise diagonals must have 7/7/3 entries")
    blob = struct.pack(
        "<22d2i", *p0, *q, *r,
        float(ta["max_gps_gap_threshold"]), float(rc["max_initial_duration"]),
        float(np.deg2rad(rts["sharp_turn_yaw_rate_threshold_deg_per_sec"])),
        float(rc["residual_threshold"]), EVAL_SKIP_SECONDS,
        int(rc["min_samples"]), int(rts["default_ekf_transition_steps_on_sharp_turn"]))
    assert len(blob) == FUSE_PARAMS_BYTES
    return np.frombuffer(blob, dtype=np.uint8).copy()


def pack_noise_grid(grid: np.ndarray, cfg=None) -> np.ndarray:
    """Vectorised pack of [H,3] = (q_xy, q_z, r) hypotheses into H FuseParams records (uint8 [H*184])."""
    base = np.frombuffer(pack_fuse_params(cfg).tobytes(), dtype=np.dtype([("d", "<f8", 22), ("i", "<i4", 2)]))
    rec = np.repeat(base, len(grid))
    d = rec["d"]
    d[:, 7] = grid[:, 0]; d[:, 8] = grid[:, 0]; d[:, 9] = grid[:, 1]          # q x, y, z
    d[:, 14] = grid[:, 2]; d[:, 15] = grid[:, 2]; d[:, 16] = grid[:, 2]       # r x, y, z
    rec["d"] = d
    return np.frombuffer(rec.tobytes(), dtype=np.uint8).copy()
