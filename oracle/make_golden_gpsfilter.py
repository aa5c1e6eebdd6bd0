"""Golden vectors for the GNSS outlier pre-filter (build container only; TEST INFRASTRUCTURE).

    python -m oracle.make_golden_gpsfilter

Runs the UNMODIFIED reference's filter_gps_outliers_ransac (EKFGPSSLAM.py:136-247: sklearn RANSACRegressor per window
and axis, unseeded) after np.random.seed(seed) on synthetic UTM tracks with gross and borderline outliers, and records the
inputs and the surviving indices.  tests/golden/gpsfilter_*.npz pin oracle.fusion_oracle.gps_filter_ransac (CPU suite:
a numpy restatement of sklearn's loop) and the drop-in's device path gsf_poly_ransac_dev (GPU suite, same seed).
"""
from __future__ import annotations

import os

import numpy as np

from . import ref_loader

HERE = os.path.dirname(os.path.abspath(__file__))
OUT_DIR = os.path.join(HERE, "..", "tests", "golden")


def make_track(seed, n, dt, outlier_frac, borderline_frac, t0=0.0):
    rng = np.random.default_rng(500 + seed)
    t = t0 + np.cumsum(np.full(n, dt) + rng.uniform(-0.2, 0.2, n) * dt)
    head = np.cumsum(rng.normal(0, 0.02, n))
    v = 9.0 + np.cumsum(rng.normal(0, 0.05, n))
    step = np.diff(t, prepend=t[0] - dt)
    xy = np.cumsum(np.stack([v * np.cos(head), v * np.sin(head)], 1) * step[:, None], axis=0)
    pos = np.column_stack([455779.0 + xy[:, 0], 5431368.0 + xy[:, 1], 112.0 + 0.5 * np.sin(t / 9.0)]) + rng.normal(0, 0.4, (n, 3))
    bad = rng.uniform(size=n) < outlier_frac
    pos[bad] += np.sign(rng.normal(size=(int(bad.sum()), 3))) * rng.uniform(25.0, 90.0, (int(bad.sum()), 3))
    edge = (~bad) & (rng.uniform(size=n) < borderline_frac)
    pos[edge, 0] += rng.choice([-1.0, 1.0], int(edge.sum())) * rng.uniform(8.0, 12.0, int(edge.sum()))      # around the 10 m threshold
    return t, pos, bad | edge


def main():
    R = ref_loader.load_reference()
    cases = [
        ("gpsfilter_sliding", 3, dict(n=420, dt=0.1, outlier_frac=0.06, borderline_frac=0.03), None),
        ("gpsfilter_sparse", 5, dict(n=90, dt=1.0, outlier_frac=0.1, borderline_frac=0.05), None),
        ("gpsfilter_global", 7, dict(n=260, dt=0.1, outlier_frac=0.08, borderline_frac=0.0), {"use_sliding_window": False}),
        ("gpsfilter_clean", 9, dict(n=300, dt=0.104, outlier_frac=0.0, borderline_frac=0.0), None),
    ]
    for name, seed, kw, override in cases:
        t, pos, planted = make_track(seed, **kw)
        cfg = dict(R.CONFIG["gps_filtering_ransac"])
        if override:
            cfg.update(override)
        np.random.seed(seed)
        with ref_loader.quiet():
            tf, pf = R.filter_gps_outliers_ransac(t, pos, cfg)
        after = np.random.random()                       # where the reference left numpy's global RNG
        kept = np.flatnonzero(np.isin(t, tf))
        assert len(kept) == len(tf) and np.array_equal(pos[kept], pf)
        np.savez_compressed(os.path.join(OUT_DIR, name + ".npz"), t=t, pos=pos, planted=planted, kept=kept, seed=np.array(seed),
                            sliding=np.array(cfg["use_sliding_window"]), rng_after=np.array(after))
        print(name, "kept", len(kept), "/", len(t), "planted", int(planted.sum()), "removed planted", int((planted & ~np.isin(np.arange(len(t)), kept)).sum()))


if __name__ == "__main__":
    main()
