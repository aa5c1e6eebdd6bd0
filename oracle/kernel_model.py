"""CPU model of the *kernel's* formulation of the fused path (test infrastructure).

The reference runs a dense 7x7 EKF step by step (EKFGPSSLAM.py:831-935).  The CUDA kernel
``gsf_fuse_batched`` (gps_optimize_slam_b200/csrc/gsf_fused.cu) computes the same outputs
with a time-parallel formulation:

  1. diagonal covariance  -> three scalar Kalman filters;
  2. telescoped odometry  -> M(q_state[i-1]) M(q_hat[i-1])^T = M(C) for a constant
     C = q_state[0] (x) conj(q_hat[0]) as long as no SLAM quaternion has zero norm, so
     x_pred[i] = x[i-1] + M(C) (p[i]-p[i-1]) and q_state[i] = C (x) q_hat[i];
  3. covariance recursion -> Moebius maps  p -> R(p+q)/(p+q+R)  composed as 2x2 matrices
     (chunked scan across the threads of a trajectory);
  4. state recursion      -> affine maps  x -> a x + b  (chunked scan);
  5. RTS over an outage   -> closed form  x_s[k] = x_f[k] + P_f[k]/P_pred[i] * (x_f[i]-x_pred[i]).

This file restates that formulation in numpy with the same chunking so the algebra (and the
scan operators with their rescaling) can be checked against oracle/fusion_oracle.py on the
CPU, before and independently of the GPU run.  It is not used by the product.
"""
from __future__ import annotations

import numpy as np

from . import fusion_oracle as fo


def _qmul(p, q):
    px, py, pz, pw = p
    qx, qy, qz, qw = q
    return np.array([
        pw * qx + qw * px + (py * qz - pz * qy),
        pw * qy + qw * py + (pz * qx - px * qz),
        pw * qz + qw * pz + (px * qy - py * qx),
        pw * qw - px * qx - py * qy - pz * qz])


def _qmat(q):
    x, y, z, w = q
    return np.array([
        [x * x - y * y - z * z + w * w, 2 * (x * y - z * w), 2 * (x * z + y * w)],
        [2 * (x * y + z * w), -x * x + y * y - z * z + w * w, 2 * (y * z - x * w)],
        [2 * (x * z - y * w), 2 * (y * z + x * w), -x * x - y * y + z * z + w * w]])


def _quat_from_matrix(m):
    tr = m[0, 0] + m[1, 1] + m[2, 2]
    choice = int(np.argmax([m[0, 0], m[1, 1], m[2, 2], tr]))
    if choice == 0:
        q = [1 - tr + 2 * m[0, 0], m[1, 0] + m[0, 1], m[2, 0] + m[0, 2], m[2, 1] - m[1, 2]]
    elif choice == 1:
        q = [m[1, 0] + m[0, 1], 1 - tr + 2 * m[1, 1], m[2, 1] + m[1, 2], m[0, 2] - m[2, 0]]
    elif choice == 2:
        q = [m[2, 0] + m[0, 2], m[2, 1] + m[1, 2], 1 - tr + 2 * m[2, 2], m[1, 0] - m[0, 1]]
    else:
        q = [m[2, 1] - m[1, 2], m[0, 2] - m[2, 0], m[1, 0] - m[0, 1], 1 + tr]
    q = np.array(q)
    return q / np.linalg.norm(q)


def yaw_zyx(q):
    x, y, z, w = q
    a, b, c, d = w - y, z - x, y + w, -x - z
    hs, hd = np.arctan2(b, a), np.arctan2(d, c)
    second = 2 * np.arctan2(np.hypot(c, d), np.hypot(a, b))
    if abs(second) <= 1e-7:
        first = 2 * hs
    elif abs(second - np.pi) <= 1e-7:
        first = -2 * hd
    else:
        first = hs - hd
    return (first + np.pi) % (2 * np.pi) - np.pi


def select_mask(ts, valid, gap, max_dur, min_pts):
    """Sim3 point selection as the kernel does it (rank / previous-valid formulation of
    EKFGPSSLAM.py:972-998).  Returns (mask, count) or (None, count) when too few."""
    n = len(ts)
    rank = np.cumsum(valid) - 1                       # rank among valid points
    nvalid = int(valid.sum())
    if nvalid < min_pts:
        return None, nvalid
    prev_t = np.full(n, np.nan)
    last = np.nan
    for i in range(n):                                # "last valid timestamp before i"
        prev_t[i] = last
        if valid[i]:
            last = ts[i]
    is_gap = valid & (rank >= 1) & (ts - prev_t > gap)
    k = int(rank[is_gap].min()) - 1 if is_gap.any() else None   # allv[:k] with k = first gap index
    first_cnt = nvalid if k is None else k
    in_first = valid & (rank < first_cnt)
    if first_cnt < min_pts:
        return valid.copy(), nvalid
    t0 = ts[valid][0]
    timed = in_first & (ts <= t0 + max_dur)
    if int(timed.sum()) < min_pts:
        return in_first, first_cnt
    return timed, int(timed.sum())


def umeyama_from_sums(src, dst, mask):
    n = int(mask.sum())
    if n < 3:
        return None
    mu_s, mu_d = src[mask].sum(0) / n, dst[mask].sum(0) / n
    a, b = src[mask] - mu_s, dst[mask] - mu_d
    H = a.T @ b
    ss = float((a * a).sum())
    U, S, Vt = np.linalg.svd(H)
    u1, u2, v1, v2 = U[:, 0], U[:, 1], Vt[0], Vt[1]
    R = np.outer(v1, u1) + np.outer(v2, u2) + np.outer(np.cross(v1, v2), np.cross(u1, u2))
    var = ss / n
    if var < 1e-12:
        s = 1.0
    else:
        s = S.sum() / (n * var)
        if s <= 1e-6:
            s = 1.0
    return R, mu_d - s * (R @ mu_s), s


def _moebius_compose(later, earlier):
    """(later o earlier) as 2x2 matrices, rescaled by a power of two so entries stay O(1)."""
    m = later @ earlier
    e = np.frexp(m.max())[1]
    return np.ldexp(m, -e)


def _moebius_apply(m, p):
    return (m[0, 0] * p + m[0, 1]) / (m[1, 0] * p + m[1, 1])


def chunk_bounds(n_items, T):
    L = -(-n_items // T)
    return [(min(t * L, n_items), min((t + 1) * L, n_items)) for t in range(T)]


def fused_model(ts, p, q, z, cfg=None, T=32, init_pose=None):
    """Whole fused path for one trajectory with pre-associated measurements ``z`` (NaN rows =
    no GNSS).  ``init_pose`` = (p0, q0) skips the Sim3 stage (EKF-only mode)."""
    cfg = fo.default_config() if cfg is None else cfg
    ekf, rts, ta, rc = cfg["ekf"], cfg["rts_decision"], cfg["time_alignment"], cfg["sim3_ransac"]
    n = len(ts)
    valid = ~np.isnan(z).any(axis=1)
    out = {}
    if init_pose is None:
        mask, cnt = select_mask(ts, valid, ta["max_gps_gap_threshold"], rc["max_initial_duration"], rc["min_samples"])
        if mask is None:
            raise ValueError("too few points")
        R, t, s = umeyama_from_sums(p, z, mask)
        qR = _quat_from_matrix(R)
        q0hat = q[0] / np.linalg.norm(q[0])
        x0 = s * (R @ p[0]) + t
        qs0 = fo._unit_or_identity(_qmul(qR, q0hat))
        out.update(R=R, t=t, s=s, sel_count=cnt)
    else:
        x0 = np.asarray(init_pose[0], float)
        qs0 = fo._unit_or_identity(np.asarray(init_pose[1], float))
        q0hat = q[0] / np.linalg.norm(q[0])
    C = _qmul(qs0, q0hat * np.array([-1, -1, -1, 1.0]))
    RC = _qmat(C)

    P0 = np.array(ekf["initial_cov_diag"][:3], float)
    Q = np.array(ekf["process_noise_diag"][:3], float)
    Rm = np.array(ekf["meas_noise_diag"], float)

    # ---- outage segments (maximal runs of invalid indices followed by a valid index)
    w_eff = np.ones(n)                 # blending weight applied to the gain at step i
    segments = []
    i = 0
    while i < n:
        if not valid[i]:
            s_ = i
            while i < n and not valid[i]:
                i += 1
            if i < n:                  # recovered at i
                e_ = i - 1
                sharp = False
                if e_ - s_ + 1 >= 2:
                    worst = 0.0
                    for a in range(s_ + 1, e_ + 1):
                        if ts[a] <= ts[a - 1]:
                            continue
                        if np.linalg.norm(q[a - 1]) == 0 or np.linalg.norm(q[a]) == 0:
                            worst = np.inf
                            break
                        y1, y2 = yaw_zyx(q[a - 1]), yaw_zyx(q[a])
                        d = np.arctan2(np.sin(y2 - y1), np.cos(y2 - y1))
                        worst = max(worst, abs(d / (ts[a] - ts[a - 1])))
                    sharp = worst > np.deg2rad(rts["sharp_turn_yaw_rate_threshold_deg_per_sec"])
                if sharp:
                    steps = rts["default_ekf_transition_steps_on_sharp_turn"]
                    if steps > 0 and 1.0 / steps < 1.0:
                        w_eff[i] = 1.0 / steps
                segments.append((s_, e_, not sharp))
        else:
            i += 1

    # ---- per-step inputs (time parallel)
    dt = np.maximum(1e-6, np.diff(ts))                     # step i uses dt[i-1]
    u = (p[1:] - p[:-1]) @ RC.T                            # u[i-1] = M(C)(p_i - p_{i-1})
    qn = q / np.linalg.norm(q, axis=1, keepdims=True)
    out_q = np.array([_qmul(C, qq) for qq in qn])
    out_q[0] = qs0

    # ---- covariance: chunked Moebius scan over steps 1..n-1
    steps = n - 1
    bounds = chunk_bounds(steps, T)
    Ppred = np.zeros((n, 3)); Pf = np.zeros((n, 3)); K = np.zeros((n, 3))
    Pf[0] = P0
    for ax in range(3):
        def elem(i):
            qq = Q[ax] * dt[i - 1]
            if valid[i]:
                return np.array([[Rm[ax], Rm[ax] * qq], [1.0, qq + Rm[ax]]])
            return np.array([[1.0, qq], [0.0, 1.0]])
        local = []
        for (b0, b1) in bounds:
            m = np.eye(2)
            for st in range(b0, b1):
                m = _moebius_compose(elem(st + 1), m)
            local.append(m)
        prefix = np.eye(2)
        for t_, (b0, b1) in enumerate(bounds):
            pcur = _moebius_apply(prefix, P0[ax])          # covariance at chunk start
            for st in range(b0, b1):
                i_ = st + 1
                pp = pcur + Q[ax] * dt[i_ - 1]
                Ppred[i_, ax] = pp
                if valid[i_]:
                    k = pp * (1.0 / (pp + Rm[ax]))
                    K[i_, ax] = k
                    pcur = (1 - k) * pp * (1 - k) + k * Rm[ax] * k
                else:
                    pcur = pp
                Pf[i_, ax] = pcur
            prefix = _moebius_compose(local[t_], prefix)

    # ---- state: chunked affine scan   x_i = a_i x_{i-1} + b_i
    Keff = K * w_eff[:, None]
    a = 1.0 - Keff[1:]
    zz = np.where(valid[1:, None], z[1:], 0.0)
    b = a * u + Keff[1:] * zz
    xf = np.zeros((n, 3)); xp = np.zeros((n, 3))
    xf[0] = x0
    loc = []
    for (b0, b1) in bounds:
        A, Bv = np.ones(3), np.zeros(3)
        for st in range(b0, b1):
            A, Bv = a[st] * A, a[st] * Bv + b[st]
        loc.append((A, Bv))
    pa, pb = np.ones(3), np.zeros(3)
    for t_, (b0, b1) in enumerate(bounds):
        x = pa * x0 + pb
        for st in range(b0, b1):
            xp[st + 1] = x + u[st]
            x = a[st] * x + b[st]
            xf[st + 1] = x
        A, Bv = loc[t_]
        pa, pb = A * pa, A * pb + Bv
    out_p = xf.copy()

    # ---- closed-form RTS over recovered outages
    for (s_, e_, do_rts) in segments:
        if not do_rts:
            continue
        i_ = e_ + 1
        delta = xf[i_] - xp[i_]
        for k in range(s_, e_ + 1):
            out_p[k] = xf[k] + (Pf[k] / Ppred[i_]) * delta
    out.update(pos=out_p, quat=out_q, segments=segments, valid=valid)
    return out
