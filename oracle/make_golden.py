"""Generate tests/golden/*.npz by running the UNMODIFIED reference (build container only).

    python -m oracle.make_golden

For each case the reference's own functions are called in the order of main_process_gui
(EKFGPSSLAM.py:958-1035) through oracle/ref_loader.py; inputs and outputs are stored so the
CPU tests can pin oracle/fusion_oracle.py and the GPU tests can pin the CUDA path without the
reference being present.  Cases:
  pairA      yolotum04.txt + 5.1Kitti04gps   (config 1 "as shipped"; lat/lon swapped by the
             loader -> zone 39N; the Sim3 cross-covariance has det < 0)
  pairB      yolotum04.txt + combined_output.txt (zone 32N, 270/271 valid, RTS over [0,1])
  synth_*    seeded synthetic trajectories (gps_optimize_slam_b200/synth.py); GNSS outages are
             expressed the way the reference sees them: missing GNSS samples (> 5 s gap).
The projection inside the reference is the Krueger stand-in of oracle/utm_kruger.py (pyproj is
not installed), so the UTM columns are "parity unpinned" by construction.
"""
from __future__ import annotations

import os
import sys

import numpy as np

from . import ref_loader

HERE = os.path.dirname(os.path.abspath(__file__))
OUT_DIR = os.path.join(HERE, "..", "tests", "golden")


def _replay(R, slam, gps, seed):
    """Steps 2-6 of main_process_gui with the reference's own functions."""
    cfg = R.CONFIG
    with ref_loader.quiet():
        aligned, valid = R.dynamic_time_alignment(slam, gps, cfg["time_alignment"])
    allv = np.where(valid)[0]
    # index selection :977-998 (orchestrator code, re-stated because main_process_gui needs dialogs)
    ts = slam["timestamps"]
    gaps = np.where(np.diff(ts[allv]) > cfg["time_alignment"]["max_gps_gap_threshold"])[0]
    first = allv[:gaps[0]] if len(gaps) else allv
    if len(first) < cfg["sim3_ransac"]["min_samples"]:
        sel = allv
    else:
        timed = first[ts[first] <= ts[first[0]] + cfg["sim3_ransac"]["max_initial_duration"]]
        sel = first if len(timed) < cfg["sim3_ransac"]["min_samples"] else timed
    rc = cfg["sim3_ransac"]
    np.random.seed(seed)
    with ref_loader.quiet():
        Rm, t, s = R.compute_sim3_transform_robust(slam["positions"][sel], aligned[sel], rc["min_samples"],
                                                   rc["residual_threshold"], rc["max_trials"], rc["min_inliers_needed"])
        Rd, td, sd = R.compute_sim3_transform(slam["positions"][sel], aligned[sel])
        sp, sq = R.transform_trajectory(slam["positions"], slam["quaternions"], Rm, t, s)
        fp, fq = R.apply_ekf_correction(slam, gps, sp, sq, cfg)
    ev = allv[ts[allv] > ts[0] + 5.0]
    from scipy.spatial import distance
    stats = []
    for traj in (slam["positions"], sp, fp):
        e = distance.cdist(traj[ev], aligned[ev], "euclidean").min(axis=1)
        stats.append([e.mean(), np.median(e), np.sqrt(np.mean(e ** 2))])
    return dict(aligned=aligned, valid=valid, sim3_indices=sel, R=Rm, t=t, s=s, R_direct=Rd, t_direct=td, s_direct=sd,
                sim3_pos=sp, sim3_quat=sq, ekf_pos=fp, ekf_quat=fq, eval_indices=ev, stats=np.array(stats))


def _save(name, **arrays):
    os.makedirs(OUT_DIR, exist_ok=True)
    path = os.path.join(OUT_DIR, name + ".npz")
    np.savez_compressed(path, **arrays)
    print(f"wrote {path} ({os.path.getsize(path) / 1024:.1f} KiB)")


def fixture_case(R, name, gps_file):
    ref = ref_loader.REFERENCE_DIR
    with ref_loader.quiet():
        slam = R.load_slam_trajectory(os.path.join(ref, "yolotum04.txt"))
        gps = R.load_gps_data(os.path.join(ref, gps_file), "GPS", R.CONFIG["gps_filtering_ransac"])
    raw = np.loadtxt(os.path.join(ref, gps_file))
    out = _replay(R, slam, gps, seed=0)
    _save(name, slam_ts=slam["timestamps"], slam_pos=slam["positions"], slam_quat=slam["quaternions"],
          gnss_raw=raw[:, :4], gps_ts=gps["timestamps"], gps_utm=gps["positions"],
          utm_zone=np.array(gps["utm_zone"]), **out)


def synth_case(R, name, seed, n=271, dt=0.104, speed=13.0, drop=(), sharp_turn_at=None, gps_stride=1):
    sys.path.insert(0, os.path.join(HERE, ".."))
    from gps_optimize_slam_b200 import synth
    tr = synth.make_trajectory(seed, n=n, dt=dt, speed=speed, sharp_turn_at=sharp_turn_at)
    keep = np.ones(n, dtype=bool)
    for a, b in drop:
        keep[a:b] = False
    keep &= (np.arange(n) % gps_stride == 0)
    slam = {"timestamps": tr["ts"], "positions": tr["pos"], "quaternions": tr["quat"]}
    gps = {"timestamps": tr["ts"][keep], "positions": tr["gps"][keep]}
    out = _replay(R, slam, gps, seed=seed)
    _save(name, slam_ts=tr["ts"], slam_pos=tr["pos"], slam_quat=tr["quat"], gps_ts=gps["timestamps"],
          gps_utm=gps["positions"], **out)


def main():
    R = ref_loader.load_reference()
    fixture_case(R, "pairA", "5.1Kitti04gps")
    fixture_case(R, "pairB", "combined_output.txt")
    synth_case(R, "synth_clean", 11)
    synth_case(R, "synth_outage", 12, drop=[(80, 150)])                       # 7.3 s gap -> outage + RTS
    synth_case(R, "synth_outage_start", 13, drop=[(0, 60), (160, 230)])       # starts without GNSS
    synth_case(R, "synth_sharp_turn", 14, drop=[(80, 150)], sharp_turn_at=110)  # sharp turn: no RTS
    synth_case(R, "synth_sparse_gnss", 15, n=400, gps_stride=3)               # real spline interpolation
    synth_case(R, "synth_long", 16, n=1000, dt=0.1, speed=10.0, drop=[(400, 470)])


if __name__ == "__main__":
    main()
