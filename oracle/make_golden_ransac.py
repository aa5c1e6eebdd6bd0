"""Golden vectors for the Sim3 RANSAC path (build container only; TEST INFRASTRUCTURE).

    python -m oracle.make_golden_ransac

Runs the UNMODIFIED reference's compute_sim3_transform_robust (EKFGPSSLAM.py:389-426) after
np.random.seed(seed) on point sets with gross outliers, and records inputs, the sample indices the
reference drew (re-drawn with the same seed: np.random.choice(n, m, replace=False) per trial, :408)
and its outputs.  tests/golden/ransac_*.npz pin oracle.fusion_oracle.sim3_ransac (CPU suite) and
gsf_sim3_ransac_dev (GPU suite).
"""
from __future__ import annotations

import os

import numpy as np

from . import ref_loader

HERE = os.path.dirname(os.path.abspath(__file__))
OUT_DIR = os.path.join(HERE, "..", "tests", "golden")


def make_case(R, name, seed, n, outlier_frac, trials, m=4, thr=4.0, min_inliers=4, collinear=False):
    rng = np.random.default_rng(1000 + seed)
    if collinear:          # KITTI-04-like: nearly straight road, metres of lateral wiggle over hundreds of metres
        s = np.linspace(0.0, 350.0, n)
        src = np.stack([s, 2.0 * np.sin(s / 60.0) + rng.normal(0, 0.05, n), 0.3 * np.cos(s / 90.0) + rng.normal(0, 0.02, n)], 1)
    else:
        src = np.cumsum(rng.normal(0, 1.0, (n, 3)) * [1.5, 1.5, 0.05], axis=0)
    from scipy.spatial.transform import Rotation
    Rt = Rotation.from_euler("zyx", [rng.uniform(-3, 3), rng.uniform(-0.05, 0.05), rng.uniform(-0.05, 0.05)]).as_matrix()
    scale = rng.uniform(0.9, 1.1)
    dst = scale * src @ Rt.T + np.array([455779.0, 5431368.0, 112.0]) + rng.normal(0, 0.15, (n, 3))
    bad = rng.uniform(size=n) < outlier_frac
    dst[bad] += rng.normal(0, 40.0, (int(bad.sum()), 3)) + np.sign(rng.normal(size=(int(bad.sum()), 3))) * 15.0
    np.random.seed(seed)
    with ref_loader.quiet():
        Rm, t, s = R.compute_sim3_transform_robust(src, dst, m, thr, trials, min_inliers)
    np.random.seed(seed)
    samples = np.stack([np.random.choice(n, m, replace=False) for _ in range(trials)]).astype(np.int32)
    ok = Rm is not None
    np.savez_compressed(os.path.join(OUT_DIR, name + ".npz"), src=src, dst=dst, samples=samples, seed=np.array(seed),
                        thr=np.array(thr), min_inliers=np.array(min_inliers), ok=np.array(ok),
                        R=Rm if ok else np.full((3, 3), np.nan), t=t if ok else np.full(3, np.nan), s=np.array(s if ok else np.nan),
                        outliers=bad)
    print(name, "ok" if ok else "None", "scale", s, "true outliers", int(bad.sum()), "/", n)


def main():
    R = ref_loader.load_reference()
    make_case(R, "ransac_outliers", seed=3, n=271, outlier_frac=0.2, trials=1000)
    make_case(R, "ransac_collinear", seed=5, n=271, outlier_frac=0.1, trials=300, collinear=True)
    make_case(R, "ransac_hopeless", seed=7, n=40, outlier_frac=1.0, trials=50, min_inliers=30)


if __name__ == "__main__":
    main()
