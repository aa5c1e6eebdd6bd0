# -*- coding: utf-8 -*-
"""Drop-in entry point: the reference's ``EKFGPSSLAM.py`` function surface on B200 kernels.

Same module name, function names, argument meaning, return conventions, ``CONFIG`` keys and
output file formats as /root/reference/EKFGPSSLAM.py; the arithmetic of every function below
runs in libgsf.so (hand-written sm_100a CUDA, gps_optimize_slam_b200/csrc) through the C ABI in
include/gsf.h.  Host code only parses files, sorts/uniques GNSS stamps and moves arrays.
There is no CPU fallback: without the built library or without a B200 every call raises.

    python EKFGPSSLAM.py SLAM.txt GNSS.txt [--truth GNSS_TRUTH.txt] [--save OUT_utm.txt]

replaces the reference's tkinter dialogs (EKFGPSSLAM.py:940-953; tkinter/matplotlib are not
part of this path).  Plotting (:470-666) is out of scope.

Differences a caller can observe, all documented in DESIGN.md:
  * ``compute_sim3_transform_robust`` runs the reference's RANSAC on the device; the sample
    indices are drawn on the host from numpy's global RNG exactly as the reference draws them,
    so a seeded run evaluates the same trials (the batched fused path keeps the all-points fit
    + residual flag, which is what RANSAC returns whenever every point is an inlier).
  * the UTM projection is the Krueger series kernel, not PROJ (pyproj is not installed).
"""
from __future__ import annotations

import sys
from typing import Any, Dict, Optional, Tuple

import numpy as np
import torch

from gps_optimize_slam_b200 import _lib, fusion
from gps_optimize_slam_b200.config import CONFIG, EVAL_SKIP_SECONDS, pack_fuse_params

_DEV = "cuda"


def _dev(a, dtype=torch.float64):
    return torch.from_numpy(np.ascontiguousarray(a)).to(device=_DEV, dtype=dtype)


def _one(n):
    return torch.tensor([0, n], dtype=torch.int64, device=_DEV)


# ----------------------------------------------------------------------------- helpers (API parity)
def calculate_relative_pose(pose1_pos, pose1_quat, pose2_pos, pose2_quat):
    """EKFGPSSLAM.py:77-92.  Kept for API compatibility: evaluated by the strict EKF kernel on a
    two-pose trajectory (delta = pose of step 1 expressed in frame 0)."""
    ts = np.array([0.0, 1.0])
    pos = np.stack([np.zeros(3), np.asarray(pose2_pos, float) - np.asarray(pose1_pos, float)])
    quat = np.stack([np.asarray(pose1_quat, float), np.asarray(pose2_quat, float)])
    z = np.full((2, 3), np.nan)
    p, q, st = fusion.ekf_strict_batched(_dev(ts), _dev(pos), _dev(quat), _dev(z), _one(2),
                                         fusion.params_tensor(CONFIG), _dev(np.zeros((1, 3))),
                                         _dev(np.array([[0.0, 0.0, 0.0, 1.0]])))
    if int(st.cpu()[0]) & _lib.ST_BAD_QUATERNION:
        return np.zeros(3), np.array([0.0, 0.0, 0.0, 1.0])
    return p[1].cpu().numpy(), q[1].cpu().numpy()


def quaternion_nlerp(q1, q2, weight_q2):
    """EKFGPSSLAM.py:94-105 (gsf_quat_nlerp_dev)."""
    out = fusion.quat_nlerp(_dev(np.asarray(q1, float)[None]), _dev(np.asarray(q2, float)[None]), _dev(np.array([float(weight_q2)])))
    return out[0].cpu().numpy()


def is_sharp_turn_in_segment(slam_quaternions_segment, slam_timestamps_segment, yaw_rate_threshold_rad_per_sec) -> bool:
    """EKFGPSSLAM.py:808-826 (gsf_sharp_turn_dev): max yaw rate over the segment against the threshold; a zero-norm
    quaternion makes the segment a sharp turn, like the reference's ValueError branch."""
    n = len(slam_quaternions_segment)
    if n < 2:
        return False
    flags, rate = fusion.sharp_turn(_dev(np.asarray(slam_timestamps_segment, float)), _dev(np.asarray(slam_quaternions_segment, float)),
                                    _one(n), float(yaw_rate_threshold_rad_per_sec))
    print(f"  pure-SLAM segment ({n} points): max yaw rate {np.rad2deg(float(rate.cpu()[0])):.2f} deg/s")
    return bool(int(flags.cpu()[0]))


def rts_smoother_segment(states_filt_segment, covs_filt_segment, states_pred_segment, covs_pred_segment):
    """EKFGPSSLAM.py:777-803 (gsf_rts_segment_dev): lists of 7-vectors / 7x7 matrices in, lists out."""
    n = len(states_filt_segment)
    if n == 0:
        return [], []
    xs, Ps = fusion.rts_segments(_dev(np.asarray(states_filt_segment, float)), _dev(np.asarray(covs_filt_segment, float)),
                                 _dev(np.asarray(states_pred_segment, float)), _dev(np.asarray(covs_pred_segment, float)), _one(n))
    xs, Ps = xs.cpu().numpy(), Ps.cpu().numpy()
    return [xs[k] for k in range(n)], [Ps[k] for k in range(n)]


class ExtendedKalmanFilter:
    """EKFGPSSLAM.py:679-772, same constructor, attributes and methods; the arithmetic of every call runs in
    gsf_ekf_step_dev (dense 7x7 covariance).  For whole trajectories use apply_ekf_correction / the batched API:
    this object costs one kernel launch per step, exactly like the reference costs one numpy round per step."""

    def __init__(self, initial_pos, initial_quat, config_params):
        initial_pos = np.asarray(initial_pos, float); initial_quat = np.asarray(initial_quat, float)
        if not (initial_pos.shape == (3,) and initial_quat.shape == (4,)):
            raise ValueError("EKF init: initial pose has the wrong shape")
        cfg = config_params
        self.state = np.concatenate([initial_pos, self.normalize_quaternion(initial_quat)]).astype(float)
        self.cov = np.diag(cfg["initial_cov_diag"]).astype(float)
        self.Q_per_sec = np.diag(cfg["process_noise_diag"]).astype(float)
        self.R = np.diag(cfg["meas_noise_diag"]).astype(float)
        if self.state.shape != (7,) or self.cov.shape != (7, 7) or self.Q_per_sec.shape != (7, 7) or self.R.shape != (3, 3):
            raise ValueError("EKF init: state / covariance / noise matrices have the wrong shape")
        self.gnss_available_prev = None
        self.gnss_update_weight = 0.0
        self.original_transition_steps = max(1, int(cfg.get("transition_steps", 10)))
        self.current_transition_steps = self.original_transition_steps
        self.weight_delta = 1.0
        self._last_predicted_state_for_blending = self.state.copy()

    @staticmethod
    def normalize_quaternion(q):
        q = np.asarray(q, float)
        out = fusion.quat_nlerp(_dev(q[None]), _dev(q[None]), _dev(np.array([0.0])))      # nlerp(q, q, 0) = q / |q|
        nrm2 = float(np.dot(q, q))
        return out[0].cpu().numpy() if nrm2 > 1e-18 else np.array([0.0, 0.0, 0.0, 1.0])

    def _step(self, mode, state, cov, motion, dt, z, w):
        dp, dq = motion
        nan3 = np.full(3, np.nan)
        o = fusion.ekf_step(mode, _dev(np.asarray(state, float)[None]), _dev(np.asarray(cov, float)[None]), _dev(np.asarray(dp, float)[None]),
                            _dev(np.asarray(dq, float)[None]), _dev(np.array([float(dt)])), _dev((nan3 if z is None else np.asarray(z, float))[None]),
                            _dev(np.diag(self.Q_per_sec).copy()), _dev(np.diag(self.R).copy()), _dev(np.array([float(w)])))
        xs, Ps, xp, Pp, fl = [t.cpu().numpy()[0] for t in o]
        if int(fl) & 2:
            raise ValueError("Found zero norm quaternions in `quat`.")          # scipy's message (Rotation.from_quat)
        return xs, Ps, xp, Pp, int(fl)

    def _predict(self, current_state, current_cov, slam_motion_update, delta_time):
        _, _, xp, Pp, _ = self._step(1, current_state, current_cov, slam_motion_update, delta_time, None, 1.0)
        return xp, Pp

    def _update(self, predicted_state, predicted_covariance, gps_pos_meas):
        gps_pos_meas = np.asarray(gps_pos_meas, float)
        if gps_pos_meas.shape != (3,) or np.isnan(gps_pos_meas).any():
            return None, None
        zero = (np.zeros(3), np.array([0.0, 0.0, 0.0, 1.0]))
        xs, Ps, _, _, fl = self._step(2, predicted_state, predicted_covariance, zero, 0.0, gps_pos_meas, 1.0)
        if not (fl & 1):
            print("warning (EKF update): innovation covariance is singular; update skipped")
            return None, None
        return xs, Ps

    def process_step(self, slam_motion_update, gps_measurement, gnss_is_available, delta_time, override_transition_steps=None):
        eff = override_transition_steps if override_transition_steps is not None else self.current_transition_steps
        self.weight_delta = 1.0 / eff if eff > 0 else 1.0
        just_recovered = gnss_is_available and (self.gnss_available_prev == False)   # noqa: E712 (None != False)
        if gnss_is_available:
            if just_recovered or eff == 0:
                self.gnss_update_weight = 1.0 if eff == 0 else self.weight_delta
            elif self.gnss_update_weight < 1.0:
                self.gnss_update_weight = min(1.0, self.gnss_update_weight + self.weight_delta)
        else:
            self.gnss_update_weight = 0.0
        use_z = gnss_is_available and gps_measurement is not None
        blend = self.gnss_update_weight if (self.gnss_update_weight < 1.0 and eff > 0) else 1.0
        xs, Ps, xp, Pp, fl = self._step(3 if use_z else 1, self.state, self.cov, slam_motion_update, delta_time,
                                        gps_measurement if use_z else None, blend)
        self._last_predicted_state_for_blending = xp.copy()
        self.state, self.cov = xs.copy(), Ps.copy()
        self.gnss_available_prev = gnss_is_available
        return self.state, self.cov, xp, Pp


# ----------------------------------------------------------------------------- loading
def _loadtxt_device(path: str, delimiter=None) -> np.ndarray:
    """np.loadtxt(path[, delimiter=...]) with the parsing on the device (gsf_parse_table_dev): the file's bytes go to the
    GPU as they are; comments, blank lines, field splitting and the decimal -> double conversion (correctly rounded, like
    strtod) run there.  Raises what numpy raises: FileNotFoundError, ValueError for a non-numeric / empty field or ragged rows."""
    raw = np.fromfile(path, dtype=np.uint8)                                   # FileNotFoundError like np.loadtxt
    if raw.size == 0:
        return np.empty((0,), dtype=np.float64)
    first = next((ln for ln in bytes(raw[:65536]).split(b"\n") if ln.split(b"#")[0].strip()), b"")
    body = first.split(b"#")[0].strip()
    ncols = len(body.split() if delimiter is None else body.split(delimiter.encode())) or 1
    table, status = fusion.parse_table(torch.from_numpy(raw).to("cuda"), 0 if delimiter is None else ord(delimiter), max_cols=ncols)
    if status & 1:
        raise ValueError("could not convert string to float")
    if status & (2 | 16):
        raise ValueError("Wrong number of columns")
    if status & 4:
        raise ValueError("a field has more than 19 significant digits with an ambiguous rounding")
    out = table.cpu().numpy()
    return out[0] if out.shape[0] == 1 else out                               # numpy squeezes a single row


def load_slam_trajectory(txt_path: str) -> Dict[str, np.ndarray]:
    """TUM file ``ts x y z qx qy qz qw`` (EKFGPSSLAM.py:110-125)."""
    try:
        table = _loadtxt_device(txt_path)
        if table.ndim == 1:
            table = table.reshape(1, -1)
        if table.shape[1] != 8:
            raise ValueError(f"SLAM file must have 8 columns (ts x y z qx qy qz qw), found {table.shape[1]}")
    except FileNotFoundError:
        raise ValueError(f"SLAM file not found: {txt_path}")
    except Exception as exc:
        raise ValueError(f"failed to load SLAM data ({txt_path}): {exc}")
    return {"timestamps": table[:, 0].astype(float), "positions": table[:, 1:4].astype(float),
            "quaternions": table[:, 4:8].astype(float)}


class UTMProjector:
    """Call-compatible with the ``pyproj.Proj`` object the reference stores in
    gps_data['projector'] (EKFGPSSLAM.py:268-270, :295): ``proj(lon, lat)`` and
    ``proj(E, N, inverse=True)``, evaluated by the Krueger-series kernels."""

    def __init__(self, zone: int, south: bool):
        self.zone, self.south = int(zone), bool(south)
        self.srs = f"+proj=utm +zone={self.zone}{' +south' if self.south else ''} +ellps=WGS84 +datum=WGS84 +units=m +no_defs"

    def __call__(self, x, y, inverse: bool = False):
        a, b = _dev(np.atleast_1d(x)), _dev(np.atleast_1d(y))
        fn = fusion.utm_inverse if inverse else fusion.utm_forward
        u, v = fn(a, b, self.zone, self.south)
        return u.cpu().numpy(), v.cpu().numpy()


def auto_utm_projection(lons: np.ndarray, lats: np.ndarray) -> Tuple[int, str]:
    """EKFGPSSLAM.py:127-134 (mean reduction on the device)."""
    if lons.size == 0 or lats.size == 0:
        raise ValueError("lon/lat arrays must not be empty")
    out = fusion.geo_zone(_dev(lons), _dev(lats)).cpu().numpy()
    return int(out[2]), (" +south" if out[3] else "")


def _window_inliers_device(t_dev, pos_dev, idx, config):
    """One window of EKFGPSSLAM.py:204-219 on the device: per axis one gsf_poly_ransac_dev fit (sklearn's RANSACRegressor
    loop).  The sample indices are drawn here exactly as sklearn draws them -- sample_without_replacement(n_window,
    min_samples) per trial from numpy's global RNG (RANSACRegressor(random_state=None)) -- and after each fit the RNG is
    rewound and advanced by the number of trials the fit really consumed, so that a seeded run of the reference and this
    function see the same random stream.  Raises ValueError where sklearn would (no consensus set)."""
    from sklearn.utils.random import sample_without_replacement
    rng = np.random.mtrand._rand
    nw, ms, max_trials = len(idx), int(config["min_samples"]), int(config["max_trials"])
    widx = torch.from_numpy(np.ascontiguousarray(idx, dtype=np.int32)).to("cuda")
    off = torch.tensor([0, nw], dtype=torch.int64, device="cuda")
    dyn = torch.from_numpy(fusion.dynamic_max_trials_table(nw, ms, max_trials)).to("cuda")
    dyn_off = torch.zeros(1, dtype=torch.int64, device="cuda")
    keep = np.ones(nw, dtype=bool)
    for axis in range(pos_dev.shape[1]):
        state = rng.get_state()
        samples = np.stack([sample_without_replacement(nw, ms, random_state=rng) for _ in range(max_trials)]).astype(np.int32)
        mask, n_trials, status = fusion.poly_ransac(t_dev, pos_dev, widx, off, torch.tensor([axis], dtype=torch.int32, device="cuda"),
                                                    torch.from_numpy(samples).to("cuda"), dyn, dyn_off, ms,
                                                    int(config["polynomial_degree"]), max_trials, float(config["residual_threshold_meters"]))
        used = int(n_trials.cpu()[0])
        rng.set_state(state)
        for _ in range(used):
            sample_without_replacement(nw, ms, random_state=rng)
        if int(status.cpu()[0]):
            raise ValueError("RANSAC could not find a valid consensus set")
        keep &= mask.cpu().numpy().astype(bool)
    return keep


def filter_gps_outliers_ransac(times, positions, config):
    """EKFGPSSLAM.py:136-247: sliding-window (or global) polynomial RANSAC per axis; a point survives if it is an
    inlier on all axes in at least one processed window.  The fits run on the device (gsf_poly_ransac_dev); the window
    enumeration below is the reference's loop (index bookkeeping)."""
    if not config.get("enabled", False) or len(times) < config["min_samples"]:
        return times, positions
    times = np.asarray(times, dtype=np.float64); positions = np.asarray(positions, dtype=np.float64)
    t_dev, pos_dev = _dev(times), _dev(positions)
    if not config.get("use_sliding_window", False):
        try:
            keep = _window_inliers_device(t_dev, pos_dev, np.arange(len(times)), config)
        except ValueError:
            return times, positions
        return times[keep], positions[keep]
    width = config["window_duration_seconds"]
    step = width * config["window_step_factor"]
    keep = np.zeros(len(times), dtype=bool)
    start, t_end = times[0], times[-1]
    while start < t_end:
        stop = start + width
        idx = np.where((times >= start) & (times < stop))[0]
        if len(idx) >= config["min_samples"]:
            try:
                keep[idx[_window_inliers_device(t_dev, pos_dev, idx, config)]] = True
            except ValueError:
                pass
        if step <= 1e-6:
            later = np.where(times > start)[0]
            if len(later) == 0:
                break
            start = times[later[0]]
        else:
            start += step
        if start >= t_end and times[-1] >= stop:
            start = max(times[0], times[-1] - width + 1e-6)
    return times[keep], positions[keep]


def load_gps_data(txt_path: str, data_label: str = "GPS",
                  filter_config_override: Optional[Dict[str, Any]] = None) -> Dict[str, Any]:
    """GNSS rows ``ts lat lon alt ...`` -> UTM (EKFGPSSLAM.py:249-289)."""
    try:
        try:
            raw = _loadtxt_device(txt_path, delimiter=" ")
        except ValueError:
            raw = _loadtxt_device(txt_path, delimiter=",")
        if raw.ndim == 1:
            raw = raw.reshape(1, -1)
        if raw.shape[1] < 4:
            raise ValueError(f"{data_label} file needs at least 4 columns (ts lat lon alt), found {raw.shape[1]}")
        ts, lat, lon, alt = raw[:, 0], raw[:, 1], raw[:, 2], raw[:, 3]
        keep = (np.abs(lat) <= 90) & (np.abs(lon) <= 180) & (lat != 0) & (lon != 0)
        if not np.all(keep):
            print(f"warning ({data_label}): dropped {int((~keep).sum())} invalid lat/lon rows")
            ts, lat, lon, alt = ts[keep], lat[keep], lon[keep], alt[keep]
            if len(ts) == 0:
                raise ValueError(f"{data_label}: no valid GNSS rows")
        zone, hemi = auto_utm_projection(lon, lat)
        projector = UTMProjector(zone, "south" in hemi)
        east, north = projector(lon, lat)
        utm = np.column_stack((east, north, alt))
        cfg = filter_config_override if filter_config_override is not None else CONFIG["gps_filtering_ransac"]
        ts_f, utm_f = filter_gps_outliers_ransac(ts, utm, cfg)
        if len(ts_f) < 2:
            raise ValueError(f"{data_label}: fewer than 2 points left after filtering")
        return {"timestamps": ts_f, "positions": utm_f,
                "utm_zone": f"{zone}{'S' if 'south' in hemi else 'N'}", "projector": projector}
    except FileNotFoundError:
        raise ValueError(f"{data_label} file not found: {txt_path}")
    except _lib.GsfError:
        raise
    except Exception as exc:
        raise ValueError(f"{data_label} processing failed: {exc}")


def utm_to_wgs84(utm_points: np.ndarray, projector) -> np.ndarray:
    """EKFGPSSLAM.py:291-296."""
    if utm_points.shape[1] != 3:
        raise ValueError("UTM points must be an Nx3 array")
    if not isinstance(projector, UTMProjector):
        raise TypeError("projector must be the object returned in gps_data['projector']")
    lon, lat = projector(utm_points[:, 0], utm_points[:, 1], inverse=True)
    return np.column_stack((lon, lat, utm_points[:, 2]))


# ----------------------------------------------------------------------------- association
def estimate_time_offset(slam_times, gps_times, max_samples: int) -> float:
    """EKFGPSSLAM.py:301-323: correlates two standardised linspaces, which are identical by
    construction, so the lag -- and the returned offset -- is 0 for every input."""
    return 0.0


def dynamic_time_alignment(slam_data, gps_data_source, time_align_config):
    """EKFGPSSLAM.py:325-387 -> (aligned [n,3] NaN-filled, valid [n] bool)."""
    slam_t = np.asarray(slam_data["timestamps"], float)
    gps_t = np.asarray(gps_data_source["timestamps"], float)
    gps_p = np.asarray(gps_data_source["positions"], float)
    n = len(slam_t)
    aligned, valid = np.full((n, 3), np.nan), np.zeros(n, dtype=bool)
    if n == 0 or len(gps_t) < 2:
        return aligned, valid
    t = gps_t + estimate_time_offset(slam_t, gps_t, time_align_config["max_samples_for_corr"])
    order = np.argsort(t)                                   # index bookkeeping stays on the host
    t, p = t[order], gps_p[order]
    tu, first = np.unique(t, return_index=True)
    if len(tu) < 2:
        return aligned, valid
    if len(tu) < len(t):
        t, p = tu, p[first]
    gap = float(time_align_config["max_gps_gap_threshold"])
    if len(t) > 8192:                                       # long tracks: local-halo spline solve, parallel over the knots
        a, v, st = fusion.associate_spline_long(_dev(t), _dev(p), _dev(slam_t), gap)
        if int(st.cpu()[0]) == 0:
            return a.cpu().numpy(), v.cpu().numpy().astype(bool)
    a, v = fusion.associate_spline(_dev(t), _dev(p), _one(len(t)), _dev(slam_t), _one(n), gap=gap)
    return a.cpu().numpy(), v.cpu().numpy().astype(bool)


# ----------------------------------------------------------------------------- Sim3
def compute_sim3_transform(src: np.ndarray, dst: np.ndarray):
    """Umeyama (EKFGPSSLAM.py:428-459) -> (R, t, s) or (None, None, None)."""
    if src.shape[0] < 3 or src.shape != dst.shape or src.shape[1] != 3:
        return None, None, None
    n = src.shape[0]
    R, t, s, st = fusion.umeyama_batched(_dev(src), _dev(dst), _one(n), n)
    if int(st.cpu()[0]) & _lib.ST_TOO_FEW_POINTS:
        return None, None, None
    return R[0].cpu().numpy(), t[0].cpu().numpy(), float(s[0].cpu())


def compute_sim3_transform_robust(src, dst, min_samples, residual_threshold, max_trials,
                                  min_inliers_needed, point_description: str = "points"):
    """EKFGPSSLAM.py:389-426 on the device.  The sample indices are drawn here exactly as the
    reference draws them -- np.random.choice(n, min_samples, replace=False) once per trial from
    numpy's global RNG (:408) -- and handed to gsf_sim3_ransac_dev, which runs every trial
    (sample fit, residuals of all points, inlier count), keeps the first strictly-best one and
    refits on its inliers.  With the same np.random.seed the reference evaluates the same trials."""
    src = np.asarray(src, dtype=np.float64); dst = np.asarray(dst, dtype=np.float64)
    n = src.shape[0]
    if n < min_samples or src.shape != dst.shape:
        return None, None, None
    samples = np.stack([np.random.choice(n, min_samples, replace=False) for _ in range(int(max_trials))]).astype(np.int32)
    R, t, s, mask, info, st = fusion.sim3_ransac(_dev(src), _dev(dst), torch.from_numpy(samples).to("cuda"),
                                                 residual_threshold, min_inliers_needed)
    info = info.cpu().numpy()
    print(f"  Sim3 RANSAC: best inlier count {int(info[0])}/{n} {point_description} (trial {int(info[1])} of {int(max_trials)})")
    if int(info[2]) == 0 or (int(st.cpu()[0]) & _lib.ST_TOO_FEW_POINTS):
        return None, None, None
    return R.cpu().numpy(), t.cpu().numpy(), float(s.cpu()[0])


def transform_trajectory(positions, quaternions, R_mat, t_vec, scale_val):
    """EKFGPSSLAM.py:461-467."""
    n = positions.shape[0]
    p, q, st = fusion.sim3_apply_batched(_dev(positions), _dev(quaternions), _one(n), n, _dev(np.asarray(R_mat)[None]),
                                         _dev(np.asarray(t_vec)[None]), _dev(np.array([scale_val], dtype=float)))
    if int(st.cpu()[0]) & _lib.ST_BAD_QUATERNION:
        raise ValueError("Found zero norm quaternions in `quat`.")
    return p.cpu().numpy(), q.cpu().numpy()


def select_sim3_indices(slam_timestamps, valid_mask, config=CONFIG):
    """Orchestrator logic of EKFGPSSLAM.py:972-998 (index bookkeeping, host)."""
    allv = np.where(valid_mask)[0]
    need = config["sim3_ransac"]["min_samples"]
    if len(allv) < need:
        raise ValueError(f"only {len(allv)} time-synchronised points for Sim3 (< {need})")
    gaps = np.where(np.diff(slam_timestamps[allv]) > config["time_alignment"]["max_gps_gap_threshold"])[0]
    first = allv[:gaps[0]] if len(gaps) else allv
    if len(first) < need:
        return allv
    timed = first[slam_timestamps[first] <= slam_timestamps[first[0]] + config["sim3_ransac"]["max_initial_duration"]]
    return first if len(timed) < need else timed


# ----------------------------------------------------------------------------- EKF
def apply_ekf_correction(slam_data_in, gps_data_in, sim3_pos_initial, sim3_quat_initial, global_config):
    """EKFGPSSLAM.py:831-935 -> (corrected_pos [n,3], corrected_quat [n,4])."""
    ts = np.asarray(slam_data_in["timestamps"], float)
    n = len(ts)
    if n == 0:
        return np.empty((0, 3)), np.empty((0, 4))
    if not (sim3_pos_initial.shape[0] == n and sim3_quat_initial.shape[0] == n):
        raise ValueError(f"Sim3-aligned trajectory has {sim3_pos_initial.shape[0]} poses, SLAM has {n}")
    aligned, valid = dynamic_time_alignment(slam_data_in, gps_data_in, global_config["time_alignment"])
    z = np.where(valid[:, None], aligned, np.nan)
    args = (_dev(ts), _dev(slam_data_in["positions"]), _dev(slam_data_in["quaternions"]), _dev(z), _one(n))
    prm = fusion.params_tensor(global_config)
    ip, iq = _dev(sim3_pos_initial[:1]), _dev(sim3_quat_initial[:1])
    p, q, _, st = fusion.fuse_batched(*args, n, prm, init_pos=ip, init_quat=iq)      # any length (tiled kernel beyond ~4000 poses)
    if int(st.cpu()[0]) & _lib.ST_BAD_QUATERNION:
        # zero-norm SLAM quaternions: the literal recursion keeps the reference's zero-motion fallback (:84-86)
        p, q, st = fusion.ekf_strict_batched(*args, prm, ip, iq)
    return p.cpu().numpy(), q.cpu().numpy()


def fuse_trajectory(slam_data, aligned, config=CONFIG):
    """Whole device path for one trajectory with associated measurements (NaN row = no GNSS): Sim3 selection +
    Umeyama + EKF in ONE kernel launch -> dict(R, t, s, pos, quat, status, sim3_path).
    The fused kernel fits ALL selected points; that equals compute_sim3_transform_robust (EKFGPSSLAM.py:389-426)
    whenever RANSAC's best trial keeps every point.  When a residual of the all-points fit reaches the threshold
    (status bit 16, GSF_ST_RANSAC_OUTLIERS) the reference would have refitted on an inlier subset, so this function
    re-runs the explicit route -- device RANSAC, transform of pose 0, EKF seeded with it -- like main_process does."""
    ts = np.asarray(slam_data["timestamps"], float)
    n = len(ts)
    args = (_dev(ts), _dev(slam_data["positions"]), _dev(slam_data["quaternions"]), _dev(aligned), _one(n))
    prm = fusion.params_tensor(config)
    p, q, sim3, st = fusion.fuse_batched(*args, n, prm)
    s3 = sim3[0].cpu().numpy()
    status = int(st.cpu()[0])
    out = {"R": s3[:9].reshape(3, 3), "t": s3[9:12], "s": float(s3[12]), "n_selected": int(s3[13]),
           "pos": p.cpu().numpy(), "quat": q.cpu().numpy(), "status": status, "sim3_path": "fused all-points fit"}
    if status & _lib.ST_RANSAC_OUTLIERS:
        valid = ~np.isnan(np.asarray(aligned)).any(axis=1)
        sel = select_sim3_indices(ts, valid, config)
        rc = config["sim3_ransac"]
        R, t, s = compute_sim3_transform_robust(np.asarray(slam_data["positions"])[sel], np.asarray(aligned)[sel], rc["min_samples"],
                                                rc["residual_threshold"], rc["max_trials"], rc["min_inliers_needed"])
        if R is None:
            raise RuntimeError("Sim3 global transform failed")
        p0, q0 = transform_trajectory(np.asarray(slam_data["positions"])[:1], np.asarray(slam_data["quaternions"])[:1], R, t, s)
        p, q, _, st2 = fusion.fuse_batched(*args, n, prm, init_pos=_dev(p0), init_quat=_dev(q0))
        out.update(R=R, t=np.asarray(t), s=float(s), n_selected=len(sel), pos=p.cpu().numpy(), quat=q.cpu().numpy(),
                   status=int(st2.cpu()[0]) | _lib.ST_RANSAC_OUTLIERS, sim3_path="RANSAC refit (outliers in the Sim3 window)")
    return out


# ----------------------------------------------------------------------------- evaluation + orchestration
def evaluate_errors(traj_xyz, aligned, slam_timestamps, skip_seconds: float = EVAL_SKIP_SECONDS):
    """EKFGPSSLAM.py:1021-1033 -> (mean, median, rmse, count) of nearest-neighbour errors."""
    n = len(slam_timestamps)
    stats = fusion.ate_nn_batched(_dev(traj_xyz), _dev(aligned), _dev(slam_timestamps), _one(n), n, skip_seconds)
    m, med, rmse, cnt = stats[0].cpu().numpy()
    if cnt < 0:
        raise _lib.GsfError(f"evaluate_errors: evaluation set of {int(-cnt)} points exceeds the kernel's staging capacity")
    return float(m), float(med), float(rmse), int(cnt)


def save_results(slam_path, out_path_utm, timestamps, corrected_pos, corrected_quat, projector):
    """Writers of EKFGPSSLAM.py:1087-1102 (same formats, headers and file naming); the text is produced on the device
    (gsf_write_pose_rows_dev: exact "%.Df" formatting) and written to disk as one block."""
    ts_d, q_d = _dev(timestamps), _dev(corrected_quat)
    utm_text = fusion.write_pose_rows(ts_d, _dev(corrected_pos), q_d, [6, 6, 6, 6, 8, 8, 8, 8], "timestamp x y z qx qy qz qw (UTM)\n")
    utm_text.cpu().numpy().tofile(out_path_utm)
    wgs = utm_to_wgs84(corrected_pos, projector)
    out_wgs = out_path_utm.replace("_utm.txt", "_wgs84.txt")
    if out_wgs == out_path_utm:
        out_wgs = out_path_utm.replace(".txt", "_wgs84.txt") if ".txt" in out_path_utm else out_path_utm + "_wgs84.txt"
    wgs_text = fusion.write_pose_rows(ts_d, _dev(wgs), q_d, [6, 8, 8, 3, 8, 8, 8, 8], "timestamp lon lat alt qx qy qz qw (WGS84)\n")
    wgs_text.cpu().numpy().tofile(out_wgs)
    return out_path_utm, out_wgs


def main_process(slam_path: str, gps_path: str, ground_truth_gps_path: str = "", save_path: str = ""):
    """Steps 1-7 of main_process_gui (EKFGPSSLAM.py:940-1110) without dialogs and plots."""
    print("step 1/7: loading data")
    slam = load_slam_trajectory(slam_path)
    gps = load_gps_data(gps_path, "primary GPS", CONFIG["gps_filtering_ransac"])
    truth = load_gps_data(ground_truth_gps_path, "GNSS truth", CONFIG["ground_truth_gps_filtering"]) if ground_truth_gps_path else None
    print(f"  SLAM poses: {len(slam['positions'])}, GNSS points: {len(gps['positions'])}, UTM zone {gps['utm_zone']}")
    if len(slam["positions"]) == 0 or len(gps["positions"]) < 2:
        raise ValueError("empty SLAM data or fewer than 2 GNSS points")
    print("step 2/7: time association")
    aligned, valid = dynamic_time_alignment(slam, gps, CONFIG["time_alignment"])
    sel = select_sim3_indices(slam["timestamps"], valid)
    print(f"  {int(valid.sum())} synchronised points, {len(sel)} used for Sim3")
    print("step 3/7: Sim3")
    rc = CONFIG["sim3_ransac"]
    R, t, s = compute_sim3_transform_robust(slam["positions"][sel], aligned[sel], rc["min_samples"], rc["residual_threshold"],
                                            rc["max_trials"], rc["min_inliers_needed"])
    if R is None:
        raise RuntimeError("Sim3 global transform failed")
    print(f"  scale = {s:.10f}")
    print("step 4/7: applying Sim3")
    sim3_pos, sim3_quat = transform_trajectory(slam["positions"], slam["quaternions"], R, t, s)
    print("step 5/7: EKF + RTS fusion")
    fused_pos, fused_quat = apply_ekf_correction(slam, gps, sim3_pos, sim3_quat, CONFIG)
    print("step 6/7: evaluation (first 5 s dropped, nearest interpolated GNSS point)")
    results = {}
    refs = [("primary GPS", aligned)]
    if truth is not None:
        refs.append(("GNSS truth", dynamic_time_alignment(slam, truth, CONFIG["time_alignment"])[0]))
    for ref_name, cand in refs:
        for label, traj in (("raw SLAM", slam["positions"]), ("Sim3 aligned", sim3_pos), ("EKF fused", fused_pos)):
            m, med, rmse, cnt = evaluate_errors(traj, cand, slam["timestamps"])
            results[(ref_name, label)] = (m, med, rmse, cnt)
            print(f"    vs {ref_name:<12} {label:<14} -> mean: {m:.3f}m, median: {med:.3f}m, RMSE: {rmse:.3f}m ({cnt} pts)")
    if save_path:
        print("step 7/7: saving")
        for path in save_results(slam_path, save_path, slam["timestamps"], fused_pos, fused_quat, gps["projector"]):
            print(f"  wrote {path}")
    return {"R": R, "t": t, "s": s, "sim3_pos": sim3_pos, "sim3_quat": sim3_quat, "ekf_pos": fused_pos,
            "ekf_quat": fused_quat, "aligned": aligned, "valid": valid, "stats": results, "utm_zone": gps["utm_zone"]}


if __name__ == "__main__":
    import argparse
    ap = argparse.ArgumentParser(description="SLAM/GNSS alignment and EKF fusion on B200 kernels")
    ap.add_argument("slam", help="TUM trajectory: ts x y z qx qy qz qw")
    ap.add_argument("gps", help="GNSS rows: ts lat lon alt ...")
    ap.add_argument("--truth", default="", help="optional GNSS ground-truth file")
    ap.add_argument("--save", default="", help="output path for the *_corrected_utm.txt file")
    a = ap.parse_args()
    try:
        main_process(a.slam, a.gps, a.truth, a.save)
    except (ValueError, RuntimeError, _lib.GsfError) as exc:
        print(f"processing failed ({type(exc).__name__}): {exc}")
        sys.exit(1)
