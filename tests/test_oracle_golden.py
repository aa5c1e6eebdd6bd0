"""Pin the oracle (oracle/fusion_oracle.py) to the golden vectors produced by the unmodified
reference (oracle/make_golden.py).  Runs anywhere; no GPU, no reference needed."""
import numpy as np
import pytest

from conftest import GOLDEN_CASES, load_golden
from oracle import fusion_oracle as fo


@pytest.mark.parametrize("case", GOLDEN_CASES)
def test_association_matches_reference(case):
    g = load_golden(case)
    aligned, valid = fo.associate(g["slam_ts"], g["gps_ts"], g["gps_utm"])
    assert np.array_equal(valid, g["valid"])
    np.testing.assert_array_equal(np.isnan(aligned), np.isnan(g["aligned"]))
    np.testing.assert_allclose(aligned[valid], g["aligned"][valid], rtol=0, atol=0)


@pytest.mark.parametrize("case", GOLDEN_CASES)
def test_sim3_selection_and_umeyama(case):
    g = load_golden(case)
    sel = fo.sim3_point_selection(g["slam_ts"], g["valid"])
    assert np.array_equal(sel, g["sim3_indices"])
    R, t, s = fo.umeyama(g["slam_pos"][sel], g["aligned"][sel])
    np.testing.assert_allclose(R, g["R_direct"], rtol=0, atol=1e-15)
    np.testing.assert_allclose(t, g["t_direct"], rtol=1e-15)
    assert abs(s - g["s_direct"]) <= 1e-15
    # the reference's RANSAC result equals the all-points fit on every golden case
    np.testing.assert_allclose(g["R"], g["R_direct"], rtol=0, atol=1e-14)


@pytest.mark.parametrize("case", GOLDEN_CASES)
def test_apply_and_ekf(case):
    g = load_golden(case)
    sp, sq = fo.sim3_apply(g["slam_pos"], g["slam_quat"], g["R"], g["t"], float(g["s"]))
    np.testing.assert_allclose(sp, g["sim3_pos"], rtol=0, atol=1e-9)
    np.testing.assert_allclose(sq, g["sim3_quat"], rtol=0, atol=1e-15)
    fp, fq = fo.ekf_fuse(g["slam_ts"], g["slam_pos"], g["slam_quat"], g["aligned"], g["valid"],
                         g["sim3_pos"][0], g["sim3_quat"][0], fo.default_config())
    np.testing.assert_allclose(fp, g["ekf_pos"], rtol=0, atol=1e-9)
    np.testing.assert_allclose(fq, g["ekf_quat"], rtol=0, atol=1e-15)


@pytest.mark.parametrize("case", GOLDEN_CASES)
def test_evaluation(case):
    g = load_golden(case)
    ev = fo.evaluation_indices(g["slam_ts"], g["valid"])
    assert np.array_equal(ev, g["eval_indices"])
    for row, traj in zip(g["stats"], (g["slam_pos"], g["sim3_pos"], g["ekf_pos"])):
        np.testing.assert_allclose(fo.error_stats(fo.nn_errors(traj, g["aligned"], ev)), row, rtol=1e-13)


def test_fixture_facts():
    """SURVEY 8c facts about the shipped pairs."""
    a, b = load_golden("pairA"), load_golden("pairB")
    assert str(a["utm_zone"]) == "39N" and str(b["utm_zone"]) == "32N"
    assert a["valid"].sum() == 271 and b["valid"].sum() == 270
    assert abs(float(a["s"]) - 0.9868243285) < 1e-9 and abs(float(b["s"]) - 0.9869858179) < 1e-9
    np.testing.assert_allclose(a["stats"][2], [0.0811, 0.0812, 0.0823], atol=5e-5)
    np.testing.assert_allclose(b["stats"][2][2], 0.0839, atol=5e-5)
    H = (a["slam_pos"] - a["slam_pos"].mean(0)).T @ (a["aligned"] - a["aligned"].mean(0))
    assert np.linalg.det(H) < 0          # pair A takes the reflection branch of Umeyama


def test_loader_and_projection_on_raw_rows():
    """load path: validity mask, zone from means, Krueger projection (parity unpinned)."""
    from oracle import utm_kruger as uk
    for case in ("pairA", "pairB"):
        g = load_golden(case)
        raw = g["gnss_raw"]
        ts, lat, lon, alt = raw[:, 0], raw[:, 1], raw[:, 2], raw[:, 3]
        keep = uk.gnss_validity_mask(lat, lon)
        zone, south = uk.utm_zone_from_means(lon[keep], lat[keep])
        assert f"{zone}{'S' if south else 'N'}" == str(g["utm_zone"])
        e, n = uk.utm_forward(lon[keep], lat[keep], zone, south)
        np.testing.assert_allclose(np.column_stack((e, n, alt[keep])), g["gps_utm"], rtol=0, atol=1e-9)


@pytest.mark.parametrize("case", ["ransac_outliers", "ransac_collinear", "ransac_hopeless"])
def test_sim3_ransac_matches_seeded_reference(case):
    """oracle.sim3_ransac against the unmodified reference's seeded compute_sim3_transform_robust
    (oracle/make_golden_ransac.py), including the sample indices the global RNG produces."""
    g = load_golden(case)
    np.random.seed(int(g["seed"]))
    draws = np.stack([np.random.choice(len(g["src"]), g["samples"].shape[1], replace=False) for _ in range(len(g["samples"]))])
    assert np.array_equal(draws, g["samples"])
    np.random.seed(int(g["seed"]))
    R, t, s = fo.sim3_ransac(g["src"], g["dst"], g["samples"].shape[1], float(g["thr"]), len(g["samples"]), int(g["min_inliers"]))
    if not bool(g["ok"]):
        assert R is None
        return
    np.testing.assert_allclose(R, g["R"], rtol=0, atol=1e-15)
    np.testing.assert_allclose(t, g["t"], rtol=1e-15)
    assert abs(s - float(g["s"])) <= 1e-15


@pytest.mark.parametrize("case", ["gpsfilter_sliding", "gpsfilter_sparse", "gpsfilter_global", "gpsfilter_clean"])
def test_gps_filter_oracle_matches_seeded_reference(case):
    """oracle.gps_filter_ransac (sklearn's RANSACRegressor loop restated in numpy) against the UNMODIFIED reference's
    filter_gps_outliers_ransac (EKFGPSSLAM.py:136-247) run with the same numpy seed: same surviving points, and numpy's
    global RNG left in the same state (i.e. every fit consumed the same number of trials)."""
    import os
    from oracle import fusion_oracle as fo
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", case + ".npz"))
    cfg = dict(fo.default_config()["gps_filtering_ransac"])
    cfg["use_sliding_window"] = bool(g["sliding"])
    np.random.seed(int(g["seed"]))
    kept = fo.gps_filter_ransac(g["t"], g["pos"], cfg)
    np.testing.assert_array_equal(kept, g["kept"])
    assert np.random.random() == float(g["rng_after"])
