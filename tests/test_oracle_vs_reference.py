"""Function-by-function comparison of the oracle with the UNMODIFIED reference, imported from
/root/reference with stubbed GUI/projection modules.  Only runs where the reference exists
(the build container); skipped elsewhere (the golden vectors cover that case)."""
import numpy as np
import pytest

from oracle import fusion_oracle as fo, ref_loader
from gps_optimize_slam_b200 import synth

pytestmark = pytest.mark.skipif(not ref_loader.reference_available(), reason="reference not present")


@pytest.fixture(scope="module")
def ref():
    return ref_loader.load_reference()


def test_config_values_match(ref):
    from gps_optimize_slam_b200.config import CONFIG
    assert CONFIG == ref.CONFIG
    for k, v in fo.DEFAULT_CONFIG.items():
        for kk, vv in v.items():
            assert ref.CONFIG[k][kk] == vv


@pytest.mark.parametrize("seed", range(5))
def test_umeyama_and_apply(ref, seed):
    tr = synth.make_trajectory(500 + seed, n=120)
    src, dst = tr["pos"], tr["gps"] if seed % 2 == 0 else tr["gps"][:, [1, 0, 2]]
    with ref_loader.quiet():
        R, t, s = ref.compute_sim3_transform(src, dst)
        sp, sq = ref.transform_trajectory(tr["pos"], tr["quat"], R, t, s)
    R2, t2, s2 = fo.umeyama(src, dst)
    np.testing.assert_array_equal(R, R2); np.testing.assert_array_equal(t, t2); assert s == s2
    sp2, sq2 = fo.sim3_apply(tr["pos"], tr["quat"], R, t, s)
    np.testing.assert_array_equal(sp, sp2); np.testing.assert_array_equal(sq, sq2)
    assert ref.compute_sim3_transform(src[:2], dst[:2]) == (None, None, None) == fo.umeyama(src[:2], dst[:2])


@pytest.mark.parametrize("case", [dict(), dict(drop=[(60, 130)]), dict(drop=[(0, 55)]),
                                  dict(drop=[(60, 130)], sharp=90), dict(drop=[(60, 130)], sharp=90, steps=4)])
def test_ekf_driver(ref, case):
    tr = synth.make_trajectory(700, n=200, sharp_turn_at=case.get("sharp"))
    keep = np.ones(200, bool)
    for a, b in case.get("drop", []):
        keep[a:b] = False
    slam = {"timestamps": tr["ts"], "positions": tr["pos"], "quaternions": tr["quat"]}
    gps = {"timestamps": tr["ts"][keep], "positions": tr["gps"][keep]}
    import copy
    cfg = copy.deepcopy(ref.CONFIG)
    cfg["rts_decision"]["default_ekf_transition_steps_on_sharp_turn"] = case.get("steps", 0)
    with ref_loader.quiet():
        aligned, valid = ref.dynamic_time_alignment(slam, gps, cfg["time_alignment"])
        R, t, s = ref.compute_sim3_transform(tr["pos"][valid], aligned[valid])
        sp, sq = ref.transform_trajectory(tr["pos"], tr["quat"], R, t, s)
        fp, fq = ref.apply_ekf_correction(slam, gps, sp, sq, cfg)
    a2, v2 = fo.associate(tr["ts"], gps["timestamps"], gps["positions"])
    np.testing.assert_array_equal(valid, v2); np.testing.assert_array_equal(aligned, a2)
    fp2, fq2 = fo.ekf_fuse(tr["ts"], tr["pos"], tr["quat"], aligned, valid, sp[0], sq[0], cfg)
    np.testing.assert_array_equal(fp, fp2); np.testing.assert_array_equal(fq, fq2)


def test_helpers(ref):
    rng = np.random.default_rng(3)
    for _ in range(50):
        p1, p2 = rng.normal(size=3), rng.normal(size=3)
        q1, q2 = rng.normal(size=4), rng.normal(size=4)
        a, b = ref.calculate_relative_pose(p1, q1, p2, q2)
        c, d = fo.relative_pose(p1, q1, p2, q2)
        np.testing.assert_array_equal(a, c); np.testing.assert_array_equal(b, d)
        w = rng.uniform(-0.2, 1.2)
        np.testing.assert_array_equal(ref.quaternion_nlerp(q1, q2, w), fo.nlerp(q1, q2, w))
    with ref_loader.quiet():
        assert ref.estimate_time_offset(np.arange(50.) * 0.1, np.arange(70.) * 0.09 + 3, 500) == 0.0
    assert fo.estimate_time_offset(np.arange(50.) * 0.1, np.arange(70.) * 0.09 + 3, 500) == 0.0
