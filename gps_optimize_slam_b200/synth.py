"""Seeded synthetic trajectories for the parity tests and the small bench configs (host, numpy).

Model: a planar vehicle with a random-walk yaw drives at constant speed
in the world (UTM) frame; the "SLAM" trajectory is the same path expressed in an arbitrary
Sim3-related frame (ground-truth scale, yaw, translation) with integrated odometry drift
(per-step position noise and yaw noise); the GNSS stream is the world path plus white noise,
already associated to the SLAM stamps (GNSS stamps = SLAM stamps, so the reference's spline
association is the identity and is stated as such).  Optional GNSS outages (NaN rows) and a
sharp turn inside an outage exercise the RTS / sharp-turn branches.

The large configs (1M x 1000) are generated on the device by ``gsf_synth_generate`` with the
same model and a counter-based RNG; tests copy a slice back to the host for the oracle.
"""
from __future__ import annotations

import numpy as np

WORLD_ORIGIN = np.array([455779.0, 5431368.0, 112.0])


def _quat_z(yaw):
    return np.stack([np.zeros_like(yaw), np.zeros_like(yaw), np.sin(yaw / 2), np.cos(yaw / 2)], axis=-1)


def _qmul(p, q):
    px, py, pz, pw = np.moveaxis(p, -1, 0)
    qx, qy, qz, qw = np.moveaxis(q, -1, 0)
    return np.stack([
        pw * qx + qw * px + (py * qz - pz * qy),
        pw * qy + qw * py + (pz * qx - px * qz),
        pw * qz + qw * pz + (px * qy - py * qx),
        pw * qw - px * qx - py * qy - pz * qz], axis=-1)


def make_trajectory(seed: int, n: int = 271, dt: float = 0.104, speed: float = 13.0,
                    outages=(), sharp_turn_at=None, quat_scale_jitter: float = 0.0):
    """One trajectory -> dict(ts[n], pos[n,3], quat[n,4], gps[n,3] (NaN = no GNSS), truth)."""
    rng = np.random.default_rng(seed)
    ts = np.arange(n) * dt + rng.uniform(0, 1e-3, n).cumsum() * 0.0
    yaw_rate = rng.normal(0.0, 0.02, n).cumsum() * 0.2
    yaw = rng.uniform(-np.pi, np.pi) + np.cumsum(yaw_rate * dt)
    if sharp_turn_at is not None:
        yaw[sharp_turn_at:] += 1.2                       # ~ 660 deg/s at dt = 0.104
    vel = speed * np.stack([np.cos(yaw), np.sin(yaw), 0.02 * np.sin(0.05 * np.arange(n))], axis=1)
    world = WORLD_ORIGIN + rng.uniform(-1e3, 1e3, 3) * np.array([1, 1, 0.01]) + np.cumsum(vel * dt, axis=0)
    world_q = _quat_z(yaw)
    # ground-truth Sim3: world = s * Rz(a) * slam + t
    s_gt, a_gt = rng.uniform(0.9, 1.1), rng.uniform(-np.pi, np.pi)
    ca, sa = np.cos(a_gt), np.sin(a_gt)
    Rg = np.array([[ca, -sa, 0], [sa, ca, 0], [0, 0, 1.0]])
    t_gt = world[0].copy()
    slam_true = ((world - t_gt) / s_gt) @ Rg                 # Rg^T (w - t)/s, row-vector form
    # odometry drift: integrated per-step position noise and yaw noise
    drift_p = np.cumsum(rng.normal(0, 0.02, (n, 3)) * np.array([1, 1, 0.2]), axis=0)
    drift_yaw = np.cumsum(rng.normal(0, np.deg2rad(0.2), n))
    slam_p = slam_true + drift_p
    slam_q = _qmul(_quat_z(np.full(n, -a_gt) + drift_yaw), world_q)
    if quat_scale_jitter:
        slam_q = slam_q * (1.0 + rng.uniform(-quat_scale_jitter, quat_scale_jitter, (n, 1)))
    sigma_g = rng.uniform(0.05, 0.5)
    gps = world + rng.normal(0, sigma_g, (n, 3))
    for (a, b) in outages:
        gps[a:b] = np.nan
    return {"ts": ts, "pos": slam_p, "quat": slam_q, "gps": gps,
            "truth": {"scale": s_gt, "yaw": a_gt, "t": t_gt, "sigma_g": sigma_g}}


def make_batch(seed0: int, batch: int, n: int, **kw):
    """Stack ``batch`` equal-length trajectories -> arrays [B,n], [B,n,3], [B,n,4], [B,n,3]."""
    items = [make_trajectory(seed0 + b, n, **kw) for b in range(batch)]
    return (np.stack([i["ts"] for i in items]), np.stack([i["pos"] for i in items]),
            np.stack([i["quat"] for i in items]), np.stack([i["gps"] for i in items]))


def make_loop_trajectory(seed: int, n: int = 4541, dt: float = 0.104, speed: float = 10.0):
    """Closed-loop track of KITTI-00 length (BASELINE config 5; no KITTI-00 data ships): one lap of a
    wobbly circle, every pose with GNSS.  Same Sim3 / drift / noise model as make_trajectory."""
    rng = np.random.default_rng(seed)
    ts = np.arange(n) * dt
    lap = 2.0 * np.pi * np.arange(n) / n
    yaw = lap + 0.15 * np.sin(5 * lap) + rng.uniform(-np.pi, np.pi)
    vel = speed * np.stack([np.cos(yaw), np.sin(yaw), 0.02 * np.sin(0.05 * np.arange(n))], axis=1)
    world = WORLD_ORIGIN + np.cumsum(vel * dt, axis=0)
    world_q = _quat_z(yaw)
    s_gt, a_gt = rng.uniform(0.9, 1.1), rng.uniform(-np.pi, np.pi)
    ca, sa = np.cos(a_gt), np.sin(a_gt)
    Rg = np.array([[ca, -sa, 0], [sa, ca, 0], [0, 0, 1.0]])
    slam_true = ((world - world[0]) / s_gt) @ Rg
    drift_p = np.cumsum(rng.normal(0, 0.02, (n, 3)) * np.array([1, 1, 0.2]), axis=0)
    drift_yaw = np.cumsum(rng.normal(0, np.deg2rad(0.2), n))
    slam_q = _qmul(_quat_z(np.full(n, -a_gt) + drift_yaw), world_q)
    gps = world + rng.normal(0, 0.3, (n, 3))
    return {"ts": ts, "pos": slam_true + drift_p, "quat": slam_q, "gps": gps}


def noise_grid(k: int = 64):
    """BASELINE config 5: k^3 log-spaced hypotheses over (Q_xy in [1e-3, 1e1], Q_z in [1e-3, 1e1],
    R in [1e-2, 1e1]), Q_quat and P0 fixed.  Returns float64 [k^3, 3] = q_xy, q_z, r."""
    qxy = np.logspace(-3, 1, k); qz = np.logspace(-3, 1, k); r = np.logspace(-2, 1, k)
    g = np.stack(np.meshgrid(qxy, qz, r, indexing="ij"), axis=-1).reshape(-1, 3)
    return g
