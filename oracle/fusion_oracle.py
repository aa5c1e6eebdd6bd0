"""numpy/scipy restatement of the reference's fusion hot path (oracle; PINNED).

Every function cites the lines of /root/reference/EKFGPSSLAM.py it follows.  The
restatement deliberately uses the same library calls the reference uses
(scipy ``Rotation``, ``interp1d(kind='cubic')``, ``np.linalg.svd/inv``,
``cdist``) and the same dense 7x7 covariance algebra, so that (a) it reproduces
the reference to the last few ulps and (b) its speed is representative of the
reference's CPU path when bench.py times it as ``cpu_baseline``.

Pinned by tests/test_oracle_vs_reference.py (build container, unmodified
reference imported through oracle/ref_loader.py) and by tests/golden/*.npz.
"""
from __future__ import annotations

import copy

import numpy as np
from scipy.interpolate import interp1d
from scipy.spatial import distance
from scipy.spatial.transform import Rotation

from . import utm_kruger

# Values of the reference's module-level CONFIG dict (EKFGPSSLAM.py:22-71).
DEFAULT_CONFIG = {
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
    "gps_filtering_ransac": {            # EKFGPSSLAM.py:39-48
        "enabled": True, "use_sliding_window": True, "window_duration_seconds": 15.0, "window_step_factor": 0.5,
        "polynomial_degree": 2, "min_samples": 6, "residual_threshold_meters": 10.0, "max_trials": 50,
    },
    "time_alignment": {"max_samples_for_corr": 500, "max_gps_gap_threshold": 5.0},
    "rts_decision": {
        "sharp_turn_yaw_rate_threshold_deg_per_sec": 45.0,
        "default_ekf_transition_steps_on_sharp_turn": 0,
    },
}


def default_config():
    return copy.deepcopy(DEFAULT_CONFIG)


# --------------------------------------------------------------------------- loaders

def load_slam_tum(path):
    """EKFGPSSLAM.py:110-125 -- 8 columns ``ts x y z qx qy qz qw``."""
    try:
        table = np.loadtxt(path)
    except Exception as exc:  # the reference folds every failure into ValueError
        raise ValueError(f"cannot load SLAM trajectory {path}: {exc}")
    table = np.atleast_2d(table)
    if table.shape[1] != 8:
        raise ValueError(f"SLAM file needs 8 columns, got {table.shape[1]}")
    return {"timestamps": table[:, 0].astype(float),
            "positions": table[:, 1:4].astype(float),
            "quaternions": table[:, 4:8].astype(float)}


def load_gnss_utm(path):
    """EKFGPSSLAM.py:249-289 without the sklearn outlier filter (out of kernel scope,
    removes nothing on the shipped data): columns ``ts lat lon alt`` (:258), validity
    mask (:259), one zone from the means (:266), z = altitude (:271)."""
    try:
        try:
            raw = np.loadtxt(path, delimiter=" ")
        except ValueError:
            raw = np.loadtxt(path, delimiter=",")
    except Exception as exc:
        raise ValueError(f"cannot load GNSS file {path}: {exc}")
    raw = np.atleast_2d(raw)
    if raw.shape[1] < 4:
        raise ValueError(f"GNSS file needs >= 4 columns, got {raw.shape[1]}")
    ts, lat, lon, alt = raw[:, 0], raw[:, 1], raw[:, 2], raw[:, 3]
    keep = utm_kruger.gnss_validity_mask(lat, lon)
    ts, lat, lon, alt = ts[keep], lat[keep], lon[keep], alt[keep]
    if ts.size == 0:
        raise ValueError("no valid GNSS rows")
    zone, south = utm_kruger.utm_zone_from_means(lon, lat)
    east, north = utm_kruger.utm_forward(lon, lat, zone, south)
    if ts.size < 2:
        raise ValueError("fewer than 2 GNSS rows")
    return {"timestamps": ts, "positions": np.column_stack((east, north, alt)),
            "utm_zone": f"{zone}{'S' if south else 'N'}", "zone": zone, "south": south}


# --------------------------------------------------------------------------- association

def estimate_time_offset(slam_t, gps_t, max_samples):
    """EKFGPSSLAM.py:301-323.  Cross-correlates two standardised linspaces; both are
    the same vector after standardisation, so the lag is always 0."""
    if len(slam_t) < 2 or len(gps_t) < 2:
        return 0.0
    m = min(max_samples, len(slam_t), len(gps_t))
    if m < 2:
        return 0.0
    a = np.linspace(slam_t.min(), slam_t.max(), m)
    b = np.linspace(gps_t.min(), gps_t.max(), m)
    an, bn = a - a.mean(), b - b.mean()
    sa, sb = an.std(), bn.std()
    if sa < 1e-9 or sb < 1e-9:
        return 0.0
    an, bn = an / sa, bn / sb
    lag = int(np.correlate(an, bn, mode="full").argmax()) - m + 1
    step = (a[-1] - a[0]) / (m - 1)
    return lag * step


def associate(slam_t, gps_t, gps_xyz, gap_threshold=5.0, max_samples=500):
    """EKFGPSSLAM.py:325-387 -> (aligned[n,3] NaN-filled, valid[n] bool)."""
    n = len(slam_t)
    aligned = np.full((n, 3), np.nan)
    valid = np.zeros(n, dtype=bool)
    if n == 0 or len(gps_t) < 2:
        return aligned, valid
    t = gps_t + estimate_time_offset(slam_t, gps_t, max_samples)
    order = np.argsort(t)
    t, xyz = t[order], gps_xyz[order]
    tu, first = np.unique(t, return_index=True)
    if len(tu) < 2:
        return aligned, valid
    if len(tu) < len(t):
        t, xyz = tu, xyz[first]
    cuts = np.where(np.diff(t) > gap_threshold)[0]
    starts = np.concatenate(([0], cuts + 1))
    ends = np.concatenate((cuts, [len(t) - 1]))
    for s, e in zip(starts, ends):
        m = e - s + 1
        if m < 2:
            continue
        ts, ps = t[s:e + 1], xyz[s:e + 1]
        if not np.all(np.diff(ts) > 1e-9):
            continue
        f = interp1d(ts, ps, axis=0, kind="cubic" if m >= 4 else "linear",
                     bounds_error=False, fill_value=np.nan)
        idx = np.where((slam_t >= ts[0] - 1e-9) & (slam_t <= ts[-1] + 1e-9))[0]
        if idx.size:
            vals = f(slam_t[idx])
            aligned[idx] = vals
            valid[idx[~np.isnan(vals).any(axis=1)]] = True
    return aligned, valid


# --------------------------------------------------------------------------- Sim3

def sim3_point_selection(slam_t, valid, gap_threshold=5.0, max_duration=180.0, min_pts=4):
    """Orchestrator logic EKFGPSSLAM.py:972-998: valid indices -> first run without
    a SLAM-time gap > threshold -> at most ``max_duration`` seconds from its start,
    with the two "too few points" fallbacks."""
    allv = np.where(valid)[0]
    if len(allv) < min_pts:
        raise ValueError("too few time-synchronised points for Sim3")
    gaps = np.where(np.diff(slam_t[allv]) > gap_threshold)[0]
    first = allv[:gaps[0]] if len(gaps) else allv
    if len(first) < min_pts:
        return allv
    timed = first[slam_t[first] <= slam_t[first[0]] + max_duration]
    return first if len(timed) < min_pts else timed


def umeyama(src, dst):
    """EKFGPSSLAM.py:428-459.  Note the scale uses det(R) *after* the reflection fix,
    i.e. s = sum(S)/sum|src_c|^2 in both branches."""
    n = src.shape[0]
    if n < 3 or src.shape != dst.shape or src.shape[1] != 3:
        return None, None, None
    mu_s, mu_d = src.mean(axis=0), dst.mean(axis=0)
    a, b = src - mu_s, dst - mu_d
    U, S, Vt = np.linalg.svd(a.T @ b)
    R = Vt.T @ U.T
    if np.linalg.det(R) < 0:
        Vt = Vt.copy()
        Vt[-1] *= -1
        R = Vt.T @ U.T
    var_s = np.sum(np.sum(a ** 2, axis=1)) / n
    tr = np.sum(S * np.diag(np.eye(3) @ np.diag([1, 1, np.linalg.det(R)])))
    if var_s < 1e-12:
        scale = 1.0
    else:
        scale = tr / (n * var_s)
        if scale <= 1e-6:
            scale = 1.0
    return R, mu_d - scale * (R @ mu_s), scale


def sim3_ransac(src, dst, min_samples, residual_threshold, max_trials, min_inliers, rng=None):
    """EKFGPSSLAM.py:389-426.  ``rng`` None -> numpy's global RNG like the reference."""
    n = src.shape[0]
    if n < min_samples or src.shape != dst.shape:
        return None, None, None
    draw = np.random.choice if rng is None else rng.choice
    best, best_mask = -1, None
    for _ in range(max_trials):
        pick = draw(n, min_samples, replace=False)
        R, t, s = umeyama(src[pick], dst[pick])
        if R is None:
            continue
        mask = np.linalg.norm(s * (src @ R.T) + t - dst, axis=1) < residual_threshold
        if mask.sum() > best:
            best, best_mask = mask.sum(), mask
    if best < min_inliers:
        return None, None, None
    return umeyama(src[best_mask], dst[best_mask])


def _dynamic_max_trials(n_inliers, n_samples, min_samples, probability=0.99):
    """sklearn/linear_model/_ransac.py:_dynamic_max_trials."""
    eps = np.spacing(1)
    ratio = n_inliers / float(n_samples)
    nom = max(eps, 1 - probability)
    denom = max(eps, 1 - ratio ** min_samples)
    if nom == 1:
        return 0
    if denom == 1:
        return float("inf")
    return abs(float(np.ceil(np.log(nom) / np.log(denom))))


def poly_ransac_fit(t, y, degree, min_samples, residual_threshold, max_trials, rng=None):
    """One RANSACRegressor fit of the GNSS pre-filter (EKFGPSSLAM.py:206-216: PolynomialFeatures(degree) +
    RANSACRegressor(min_samples, residual_threshold, max_trials)), sklearn's loop restated in numpy: absolute loss,
    inliers = residual <= threshold, skip on fewer inliers, R^2 tie-break, dynamic max_trials.  The polynomial is fitted
    in centred, scaled time (same polynomial space as sklearn's Vandermonde columns).  Draws the samples like sklearn:
    sample_without_replacement from numpy's global RNG (``rng`` None).  Returns (inlier_mask, n_trials); raises
    ValueError when no trial has an inlier (sklearn does)."""
    from sklearn.utils.random import sample_without_replacement
    rs = np.random.mtrand._rand if rng is None else rng
    n = len(t)
    n_best, score_best, mask_best = 1, -np.inf, None
    trials, limit = 0, max_trials
    while trials < limit:
        trials += 1
        pick = sample_without_replacement(n, min_samples, random_state=rs)
        tm = t[pick].mean()
        sc = np.abs(t[pick] - tm).max() or 1.0
        u = (t[pick] - tm) / sc
        V = np.vander(u, degree + 1, increasing=True)
        coef, *_ = np.linalg.lstsq(V, y[pick] - y[pick][0], rcond=None)
        pred = np.vander((t - tm) / sc, degree + 1, increasing=True) @ coef + y[pick][0]
        mask = np.abs(y - pred) <= residual_threshold
        cnt = int(mask.sum())
        if cnt < n_best:
            continue
        yi = y[mask]
        sst = ((yi - yi.mean()) ** 2).sum()
        ssr = ((yi - pred[mask]) ** 2).sum()
        score = 1.0 - ssr / sst if sst > 0 else (1.0 if ssr == 0 else 0.0)
        if cnt == n_best and score < score_best:
            continue
        n_best, score_best, mask_best = cnt, score, mask
        limit = min(limit, _dynamic_max_trials(n_best, n, min_samples))
    if mask_best is None:
        raise ValueError("RANSAC could not find a valid consensus set")
    return mask_best, trials


def gps_filter_ransac(times, positions, cfg, rng=None):
    """filter_gps_outliers_ransac (EKFGPSSLAM.py:136-247): windows, per-axis fits, AND over the axes, OR over the
    windows.  -> kept indices."""
    n = len(times)
    if not cfg.get("enabled", False) or n < cfg["min_samples"]:
        return np.arange(n)

    def inliers(idx):
        keep = np.ones(len(idx), dtype=bool)
        for axis in range(positions.shape[1]):
            m, _ = poly_ransac_fit(times[idx], positions[idx, axis], cfg["polynomial_degree"], cfg["min_samples"],
                                   cfg["residual_threshold_meters"], cfg["max_trials"], rng)
            keep &= m
        return keep

    if not cfg.get("use_sliding_window", False):
        try:
            return np.flatnonzero(inliers(np.arange(n)))
        except ValueError:
            return np.arange(n)
    width = cfg["window_duration_seconds"]
    step = width * cfg["window_step_factor"]
    keep = np.zeros(n, dtype=bool)
    start, t_end = times[0], times[-1]
    while start < t_end:
        stop = start + width
        idx = np.where((times >= start) & (times < stop))[0]
        if len(idx) >= cfg["min_samples"]:
            try:
                keep[idx[inliers(idx)]] = True
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
    return np.flatnonzero(keep)


def sim3_apply(pos, quat, R, t, s):
    """EKFGPSSLAM.py:461-467."""
    rot = Rotation.from_matrix(R)
    out_q = np.array([(rot * Rotation.from_quat(q)).as_quat() for q in quat])
    return s * (pos @ R.T) + t, out_q


# --------------------------------------------------------------------------- EKF

def _unit_or_identity(q):
    """ExtendedKalmanFilter.normalize_quaternion, EKFGPSSLAM.py:697-700."""
    nrm = np.linalg.norm(q)
    return q / nrm if nrm > 1e-9 else np.array([0.0, 0.0, 0.0, 1.0])


def relative_pose(p1, q1, p2, q2):
    """EKFGPSSLAM.py:77-92."""
    try:
        r1_inv = Rotation.from_quat(q1).inv()
        r2 = Rotation.from_quat(q2)
    except ValueError:
        return np.zeros(3), np.array([0.0, 0.0, 0.0, 1.0])
    return r1_inv.apply(p2 - p1), (r1_inv * r2).as_quat()


def nlerp(q1, q2, w2):
    """EKFGPSSLAM.py:94-105."""
    if np.dot(q1, q2) < 0.0:
        q2 = -q2
    w = float(np.clip(w2, 0.0, 1.0))
    q = (1.0 - w) * q1 + w * q2
    nrm = np.linalg.norm(q)
    if nrm < 1e-9:
        return q1 if w2 < 0.5 else q2
    return q / nrm


class DenseEKF:
    """ExtendedKalmanFilter, EKFGPSSLAM.py:679-772 (dense 7x7 algebra kept)."""

    def __init__(self, p0, q0, ekf_cfg):
        self.x = np.concatenate([p0, _unit_or_identity(q0)]).astype(float)
        self.P = np.diag(ekf_cfg["initial_cov_diag"]).astype(float)
        self.Q = np.diag(ekf_cfg["process_noise_diag"]).astype(float)
        self.R = np.diag(ekf_cfg["meas_noise_diag"]).astype(float)
        self.prev_avail = None
        self.weight = 0.0
        self.steps = max(1, int(ekf_cfg.get("transition_steps", 10)))

    def predict(self, motion, dt):
        """:702-715."""
        dp, dq = motion
        rot = Rotation.from_quat(self.x[3:])
        pos = self.x[:3] + rot.apply(dp)
        q = _unit_or_identity((rot * Rotation.from_quat(dq)).as_quat())
        P = self.P + self.Q * max(abs(dt), 1e-6)
        return np.concatenate([pos, q]), (P + P.T) / 2.0

    def update(self, xp, Pp, z):
        """:717-734."""
        if z.shape != (3,) or np.isnan(z).any():
            return None, None
        H = np.zeros((3, 7))
        H[0, 0] = H[1, 1] = H[2, 2] = 1.0
        S = H @ Pp @ H.T + self.R
        S = (S + S.T) / 2.0
        try:
            S_inv = np.linalg.inv(S)
        except np.linalg.LinAlgError:
            S_inv = np.linalg.pinv(S)
        K = Pp @ H.T @ S_inv
        x = xp + K @ (z - xp[:3])
        x[3:] = _unit_or_identity(x[3:])
        IKH = np.eye(7) - K @ H
        P = IKH @ Pp @ IKH.T + K @ self.R @ K.T
        return x, (P + P.T) / 2.0

    def step(self, motion, z, avail, dt, transition_steps):
        """process_step, :736-772."""
        eff = transition_steps
        wdelta = 1.0 / eff if eff > 0 else 1.0
        xp, Pp = self.predict(motion, dt)
        xu = Pu = None
        if avail and z is not None:
            xu, Pu = self.update(xp, Pp, z)
        recovered = avail and (self.prev_avail == False)  # noqa: E712 (None != False)
        if avail:
            if recovered or eff == 0:
                self.weight = 1.0 if eff == 0 else wdelta
            elif self.weight < 1.0:
                self.weight = min(1.0, self.weight + wdelta)
        else:
            self.weight = 0.0
        x, P = xp, Pp
        if avail and xu is not None:
            if self.weight < 1.0 and eff > 0:
                w = self.weight
                x = np.concatenate([(1.0 - w) * xp[:3] + w * xu[:3], nlerp(xp[3:], xu[3:], w)])
                P = Pu
            else:
                x, P = xu, Pu
        self.x, self.P = x.copy(), P.copy()
        self.prev_avail = avail
        return self.x, self.P, xp, Pp


def rts_segment(xf, Pf, xp, Pp):
    """EKFGPSSLAM.py:777-803 (F = I); returns smoothed states only (covariances are
    discarded by the caller, :917)."""
    m = len(xf)
    xs = [None] * m
    Ps = [None] * m
    if m == 0:
        return xs
    xs[-1], Ps[-1] = xf[-1].copy(), Pf[-1].copy()
    for k in range(m - 2, -1, -1):
        try:
            A = Pf[k] @ np.linalg.inv(Pp[k + 1])
        except np.linalg.LinAlgError:
            A = Pf[k] @ np.linalg.pinv(Pp[k + 1])
        xs[k] = xf[k] + A @ (xs[k + 1] - xp[k + 1])
        xs[k][3:] = _unit_or_identity(xs[k][3:])
        Pk = Pf[k] + A @ (Ps[k + 1] - Pp[k + 1]) @ A.T
        Ps[k] = (Pk + Pk.T) / 2.0
    return xs


def sharp_turn(quats, times, thresh_rad_s):
    """EKFGPSSLAM.py:808-826."""
    if len(quats) < 2:
        return False
    worst = 0.0
    for a in range(1, len(quats)):
        t1, t2 = times[a - 1], times[a]
        if t2 <= t1:
            continue
        try:
            y1 = Rotation.from_quat(quats[a - 1]).as_euler("zyx")[0]
            y2 = Rotation.from_quat(quats[a]).as_euler("zyx")[0]
        except ValueError:
            return True
        d = np.arctan2(np.sin(y2 - y1), np.cos(y2 - y1))
        worst = max(worst, abs(d / (t2 - t1)))
    return worst > thresh_rad_s


def ekf_fuse(slam_t, slam_p, slam_q, aligned, valid, p_init, q_init, cfg, return_events=False):
    """apply_ekf_correction, EKFGPSSLAM.py:831-935, taking the association result
    (``aligned``/``valid``) as input instead of recomputing it (:847 is a pure
    function of the same arguments).  ``p_init``/``q_init`` are row 0 of the
    Sim3-aligned trajectory (:842)."""
    import warnings
    n = len(slam_t)
    if n == 0:
        return np.empty((0, 3)), np.empty((0, 4))
    ekf_cfg, rts_cfg = cfg["ekf"], cfg["rts_decision"]
    f = DenseEKF(np.asarray(p_init, float), np.asarray(q_init, float), ekf_cfg)
    f.prev_avail = bool(valid[0])
    hist_xf, hist_Pf = [f.x.copy()], [f.P.copy()]
    hist_xp, hist_Pp = [f.x.copy()], [f.P.copy()]
    out_p, out_q = np.zeros((n, 3)), np.zeros((n, 4))
    out_p[0], out_q[0] = f.x[:3], f.x[3:]
    outage = not f.prev_avail
    start = 0 if outage else -1
    events = []
    t_last = slam_t[0]
    for i in range(1, n):
        dt = max(1e-6, slam_t[i] - t_last)
        motion = relative_pose(slam_p[i - 1], slam_q[i - 1], slam_p[i], slam_q[i])
        avail = bool(valid[i])
        z = aligned[i] if avail and not np.isnan(aligned[i]).any() else None
        if z is None:
            avail = False
        do_rts, steps_here = True, 0
        if not avail and not outage:
            outage, start = True, i
        elif avail and outage:
            idx = range(start, i)
            if len(idx) >= 2:
                with warnings.catch_warnings():
                    warnings.simplefilter("ignore")
                    turn = sharp_turn([slam_q[k] for k in idx], [slam_t[k] for k in idx],
                                      np.deg2rad(rts_cfg["sharp_turn_yaw_rate_threshold_deg_per_sec"]))
                if turn:
                    do_rts = False
                    steps_here = rts_cfg["default_ekf_transition_steps_on_sharp_turn"]
        xf, Pf, xp, Pp = f.step(motion, z, avail, dt, steps_here if (avail and outage) else 0)
        hist_xf.append(xf.copy()); hist_Pf.append(Pf.copy())
        hist_xp.append(xp.copy()); hist_Pp.append(Pp.copy())
        out_p[i], out_q[i] = xf[:3], xf[3:]
        if avail and outage:
            if do_rts and i - start + 1 > 1:
                sm = rts_segment(hist_xf[start:i + 1], hist_Pf[start:i + 1],
                                 hist_xp[start:i + 1], hist_Pp[start:i + 1])
                for k, xs in enumerate(sm):
                    out_p[start + k], out_q[start + k] = xs[:3], xs[3:]
                    hist_xf[start + k] = xs.copy()
            events.append((start, i, bool(do_rts)))
            outage, start = False, -1
        t_last = slam_t[i]
    if return_events:
        return out_p, out_q, events
    return out_p, out_q


# --------------------------------------------------------------------------- evaluation

def evaluation_indices(slam_t, valid, skip_seconds=5.0):
    """EKFGPSSLAM.py:1015-1023."""
    idx = np.where(valid)[0]
    if idx.size == 0:
        return idx
    return idx[slam_t[idx] > slam_t[0] + skip_seconds]


def nn_errors(traj_xyz, aligned, idx):
    """EKFGPSSLAM.py:1024-1033: nearest-neighbour distance to the candidate set."""
    if idx.size == 0:
        return np.empty(0)
    return distance.cdist(traj_xyz[idx], aligned[idx], "euclidean").min(axis=1)


def error_stats(err):
    """mean, median, RMSE as printed at EKFGPSSLAM.py:1033."""
    if err.size == 0:
        return np.array([np.nan, np.nan, np.nan])
    return np.array([err.mean(), np.median(err), np.sqrt(np.mean(err ** 2))])


# --------------------------------------------------------------------------- whole path

def run_pipeline(slam, gps, cfg=None, ransac_seed=None, use_ransac=True):
    """Steps 2-6 of main_process_gui, EKFGPSSLAM.py:970-1035, on loaded data."""
    cfg = default_config() if cfg is None else cfg
    ta, rc = cfg["time_alignment"], cfg["sim3_ransac"]
    t, p, q = slam["timestamps"], slam["positions"], slam["quaternions"]
    aligned, valid = associate(t, gps["timestamps"], gps["positions"],
                               ta["max_gps_gap_threshold"], ta["max_samples_for_corr"])
    sel = sim3_point_selection(t, valid, ta["max_gps_gap_threshold"],
                               rc["max_initial_duration"], rc["min_samples"])
    if use_ransac:
        rng = None if ransac_seed is None else np.random.default_rng(ransac_seed)
        R, tr, s = sim3_ransac(p[sel], aligned[sel], rc["min_samples"], rc["residual_threshold"],
                               rc["max_trials"], rc["min_inliers_needed"], rng)
    else:
        R, tr, s = umeyama(p[sel], aligned[sel])
    if R is None:
        raise RuntimeError("Sim3 failed")
    sp, sq = sim3_apply(p, q, R, tr, s)
    fp, fq, events = ekf_fuse(t, p, q, aligned, valid, sp[0], sq[0], cfg, return_events=True)
    ev = evaluation_indices(t, valid)
    stats = {name: error_stats(nn_errors(traj, aligned, ev))
             for name, traj in (("raw", p), ("sim3", sp), ("ekf", fp))}
    return {"aligned": aligned, "valid": valid, "sim3_indices": sel, "R": R, "t": tr, "s": s,
            "sim3_pos": sp, "sim3_quat": sq, "ekf_pos": fp, "ekf_quat": fq,
            "eval_indices": ev, "stats": stats, "rts_events": events}
