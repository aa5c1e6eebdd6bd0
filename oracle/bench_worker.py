"""CPU-baseline worker: runs the oracle's restatement of the reference pipeline (Sim3 point
selection -> Umeyama -> transform -> dense-7x7 EKF/RTS) on a list of trajectories.  Used only by
bench.py's ``cpu_baseline`` leg and ``--impl reference`` arm, in spawned worker processes."""
from __future__ import annotations

import time

import numpy as np


def run_trajectories(batch):
    """batch: list of (ts, pos, quat, z).  Returns (pose_updates, seconds)."""
    from . import fusion_oracle as fo
    cfg = fo.default_config()
    updates = 0
    t0 = time.perf_counter()
    for ts, p, q, z in batch:
        valid = ~np.isnan(z).any(axis=1)
        sel = fo.sim3_point_selection(ts, valid)
        R, t, s = fo.umeyama(p[sel], z[sel])
        sp, sq = fo.sim3_apply(p, q, R, t, s)
        fo.ekf_fuse(ts, p, q, z, valid, sp[0], sq[0], cfg)
        updates += len(ts) - 1
    return updates, time.perf_counter() - t0


def warm():
    import scipy.spatial.transform  # noqa: F401  (import cost outside the timed region)
    from . import fusion_oracle  # noqa: F401
    return True
