"""Parity of the CUDA path (through the C ABI) with the oracle and the golden vectors.
Run on a B200:  python -m pytest tests -m gpu -x -q

Tolerances (north_star: 1e-9 relative in fp64).  UTM-scale coordinates are ~5.4e6 m, where
one fp64 ulp is 9.3e-10 m; positions are compared with an absolute tolerance of 2e-7 m
(4e-14 relative), quaternions / rotations with 1e-9 absolute."""
import os

import numpy as np
import pytest

from conftest import GOLDEN_CASES, load_golden

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

POS_ATOL = 2e-7
ROT_ATOL = 1e-9


@pytest.fixture(scope="module")
def gsf():
    if not torch.cuda.is_available():
        pytest.fail("GPU tests need a CUDA device")
    from gps_optimize_slam_b200 import _lib, fusion
    _lib.load()                                   # raises when libgsf.so is missing
    return fusion


def dev(a, dtype=torch.float64):
    return torch.from_numpy(np.ascontiguousarray(a)).to("cuda", dtype=dtype)


def pack(trajs):
    """list of dict(ts,pos,quat,gps) -> flat device arrays + offsets."""
    lens = [len(t["ts"]) for t in trajs]
    off = np.concatenate(([0], np.cumsum(lens))).astype(np.int64)
    cat = lambda k: np.concatenate([t[k] for t in trajs])
    return dev(cat("ts")), dev(cat("pos")), dev(cat("quat")), dev(cat("gps")), dev(off, torch.int64), off, max(lens)


def oracle_pipeline(tr, cfg):
    from oracle import fusion_oracle as fo
    ts, p, q, z = tr["ts"], tr["pos"], tr["quat"], tr["gps"]
    valid = ~np.isnan(z).any(1)
    sel = fo.sim3_point_selection(ts, valid, cfg["time_alignment"]["max_gps_gap_threshold"],
                                  cfg["sim3_ransac"]["max_initial_duration"], cfg["sim3_ransac"]["min_samples"])
    R, t, s = fo.umeyama(p[sel], z[sel])
    sp, sq = fo.sim3_apply(p, q, R, t, s)
    fp, fq = fo.ekf_fuse(ts, p, q, z, valid, sp[0], sq[0], cfg)
    return dict(R=R, t=t, s=s, sel=sel, sim3_pos=sp, sim3_quat=sq, pos=fp, quat=fq, valid=valid)


def mixed_batch():
    from gps_optimize_slam_b200 import synth
    specs = [dict(seed=1, n=271), dict(seed=2, n=271, outages=[(50, 90)]), dict(seed=3, n=150, outages=[(0, 9), (70, 71)]),
             dict(seed=4, n=271, outages=[(50, 90)], sharp_turn_at=70), dict(seed=5, n=1000, dt=0.1, speed=10.0, outages=[(300, 420)]),
             dict(seed=6, n=33), dict(seed=7, n=400, outages=[(100, 160)]), dict(seed=8, n=271, outages=[(250, 271)]),
             dict(seed=9, n=64, quat_scale_jitter=1e-3), dict(seed=10, n=5), dict(seed=11, n=97, outages=[(10, 12), (20, 23), (40, 41)]),
             dict(seed=12, n=272)]
    return [synth.make_trajectory(**s) for s in specs]


def test_device_present(gsf):
    from gps_optimize_slam_b200 import _lib
    assert _lib.load().gsf_device_sm_count() > 0


@pytest.mark.parametrize("steps", [0, 4])
def test_fused_ragged_batch_matches_oracle(gsf, steps):
    from oracle import fusion_oracle as fo
    cfg = fo.default_config()
    cfg["rts_decision"]["default_ekf_transition_steps_on_sharp_turn"] = steps
    trajs = mixed_batch()
    ts, pos, quat, z, off_d, off, max_len = pack(trajs)
    p, q, sim3, st = gsf.fuse_batched(ts, pos, quat, z, off_d, max_len, gsf.params_tensor(cfg))
    p, q, sim3, st = p.cpu().numpy(), q.cpu().numpy(), sim3.cpu().numpy(), st.cpu().numpy()
    for b, tr in enumerate(trajs):
        o = oracle_pipeline(tr, cfg)
        sl = slice(off[b], off[b + 1])
        assert st[b] == 0, (b, st[b])
        assert int(sim3[b, 13]) == len(o["sel"]) and int(sim3[b, 14]) == int(o["valid"].sum())
        np.testing.assert_allclose(sim3[b, :9].reshape(3, 3), o["R"], atol=ROT_ATOL, err_msg=f"R traj {b}")
        np.testing.assert_allclose(sim3[b, 9:12], o["t"], rtol=1e-12, atol=POS_ATOL)
        assert abs(sim3[b, 12] - o["s"]) < 1e-11
        np.testing.assert_allclose(p[sl], o["pos"], rtol=0, atol=POS_ATOL, err_msg=f"pos traj {b}")
        np.testing.assert_allclose(q[sl], o["quat"], rtol=0, atol=ROT_ATOL, err_msg=f"quat traj {b}")


@pytest.mark.parametrize("case", GOLDEN_CASES)
def test_fused_matches_reference_golden(gsf, case):
    """Inputs and expected outputs produced by the unmodified reference (tests/golden)."""
    g = load_golden(case)
    n = len(g["slam_ts"])
    off = dev(np.array([0, n]), torch.int64)
    p, q, sim3, st = gsf.fuse_batched(dev(g["slam_ts"]), dev(g["slam_pos"]), dev(g["slam_quat"]), dev(g["aligned"]), off, n,
                                      gsf.params_tensor())
    sim3 = sim3.cpu().numpy()[0]
    assert int(st.cpu()[0]) == 0
    assert int(sim3[13]) == len(g["sim3_indices"])
    np.testing.assert_allclose(sim3[:9].reshape(3, 3), g["R"], atol=ROT_ATOL)
    np.testing.assert_allclose(sim3[9:12], g["t"], rtol=1e-12, atol=POS_ATOL)
    assert abs(sim3[12] - float(g["s"])) < 1e-11
    np.testing.assert_allclose(p.cpu().numpy(), g["ekf_pos"], rtol=0, atol=POS_ATOL)
    np.testing.assert_allclose(q.cpu().numpy(), g["ekf_quat"], rtol=0, atol=ROT_ATOL)


def test_strict_kernel_and_ekf_only_mode(gsf):
    from oracle import fusion_oracle as fo
    cfg = fo.default_config()
    trajs = mixed_batch()
    ts, pos, quat, z, off_d, off, max_len = pack(trajs)
    outs = [oracle_pipeline(t, cfg) for t in trajs]
    ip = dev(np.stack([o["sim3_pos"][0] for o in outs]))
    iq = dev(np.stack([o["sim3_quat"][0] for o in outs]))
    prm = gsf.params_tensor(cfg)
    ps, qs, st = gsf.ekf_strict_batched(ts, pos, quat, z, off_d, prm, ip, iq)
    pf, qf, _, st2 = gsf.fuse_batched(ts, pos, quat, z, off_d, max_len, prm, init_pos=ip, init_quat=iq)
    assert not st.cpu().numpy().any() and not st2.cpu().numpy().any()
    exp_p = np.concatenate([o["pos"] for o in outs]); exp_q = np.concatenate([o["quat"] for o in outs])
    for got_p, got_q in ((ps, qs), (pf, qf)):
        np.testing.assert_allclose(got_p.cpu().numpy(), exp_p, rtol=0, atol=POS_ATOL)
        np.testing.assert_allclose(got_q.cpu().numpy(), exp_q, rtol=0, atol=ROT_ATOL)


def test_strict_kernel_zero_quaternion_fallback(gsf):
    from gps_optimize_slam_b200 import synth, _lib
    from oracle import fusion_oracle as fo
    cfg = fo.default_config()
    tr = synth.make_trajectory(77, outages=[(30, 40)])
    o = oracle_pipeline(tr, cfg)
    tr["quat"][100] = 0; tr["quat"][35] = 0; tr["quat"][36] = 0
    fp, fq = fo.ekf_fuse(tr["ts"], tr["pos"], tr["quat"], tr["gps"], o["valid"], o["sim3_pos"][0], o["sim3_quat"][0], cfg)
    ts, pos, quat, z, off_d, off, max_len = pack([tr])
    prm = gsf.params_tensor(cfg)
    ip, iq = dev(o["sim3_pos"][:1]), dev(o["sim3_quat"][:1])
    ps, qs, st = gsf.ekf_strict_batched(ts, pos, quat, z, off_d, prm, ip, iq)
    assert int(st.cpu()[0]) == _lib.ST_BAD_QUATERNION
    np.testing.assert_allclose(ps.cpu().numpy(), fp, rtol=0, atol=POS_ATOL)
    np.testing.assert_allclose(qs.cpu().numpy(), fq, rtol=0, atol=ROT_ATOL)
    # the fused pipeline mirrors the reference run, which aborts (scipy ValueError at :466)
    _, _, _, stf = gsf.fuse_batched(ts, pos, quat, z, off_d, max_len, prm)
    assert int(stf.cpu()[0]) & _lib.ST_BAD_QUATERNION


def test_edge_cases_status(gsf):
    from gps_optimize_slam_b200 import synth, _lib
    a = synth.make_trajectory(21, n=3)                 # < min_samples(4) points
    b = synth.make_trajectory(22, n=40); b["gps"][:] = np.nan      # no GNSS at all
    c = synth.make_trajectory(23, n=1)
    d = synth.make_trajectory(24, n=50)
    e = synth.make_trajectory(25, n=50); e["gps"][3:] = np.nan     # 3 valid points only
    ts, pos, quat, z, off_d, off, max_len = pack([a, b, c, d, e])
    p, q, sim3, st = gsf.fuse_batched(ts, pos, quat, z, off_d, max_len, gsf.params_tensor())
    st = st.cpu().numpy(); p = p.cpu().numpy()
    assert (st[[0, 1, 2, 4]] & _lib.ST_TOO_FEW_POINTS).all() and st[3] == 0
    assert np.isnan(p[off[0]:off[3]]).all() and np.isfinite(p[off[3]:off[4]]).all()


def test_umeyama_and_apply_kernels(gsf):
    from oracle import fusion_oracle as fo
    trajs = mixed_batch()
    src = np.concatenate([t["pos"] for t in trajs]); dst = np.concatenate([np.nan_to_num(t["gps"], nan=1.0) for t in trajs])
    dst[: len(trajs[0]["ts"])] = dst[: len(trajs[0]["ts"])][:, [1, 0, 2]]       # reflection branch for trajectory 0
    lens = [len(t["ts"]) for t in trajs]
    off = np.concatenate(([0], np.cumsum(lens))).astype(np.int64)
    R, t, s, st = gsf.umeyama_batched(dev(src), dev(dst), dev(off, torch.int64), max(lens))
    quat = np.concatenate([tr["quat"] for tr in trajs])
    # the apply kernel is checked with the oracle's own transforms (same R on both sides)
    fits = [fo.umeyama(src[off[b]:off[b + 1]], dst[off[b]:off[b + 1]]) for b in range(len(trajs))]
    ap, aq, ast = gsf.sim3_apply_batched(dev(src), dev(quat), dev(off, torch.int64), max(lens), dev(np.stack([f[0] for f in fits])),
                                         dev(np.stack([f[1] for f in fits])), dev(np.array([f[2] for f in fits])))
    R, t, s, ap, aq = [x.cpu().numpy() for x in (R, t, s, ap, aq)]
    for b in range(len(trajs)):
        sl = slice(off[b], off[b + 1])
        Ro, to, so = fits[b]
        np.testing.assert_allclose(R[b], Ro, atol=ROT_ATOL); np.testing.assert_allclose(t[b], to, rtol=1e-11, atol=POS_ATOL)
        assert abs(s[b] - so) < 1e-11 * max(1.0, abs(so))
        po, qo = fo.sim3_apply(src[sl], quat[sl], Ro, to, so)
        np.testing.assert_allclose(ap[sl], po, rtol=0, atol=POS_ATOL); np.testing.assert_allclose(aq[sl], qo, rtol=0, atol=ROT_ATOL)
    # large single trajectory spanning several reduction tiles + masked points
    rng = np.random.default_rng(5)
    n = 50000
    big = rng.normal(size=(n, 3)) * [300, 200, 5]
    Rt = fo.Rotation.random(random_state=3).as_matrix()
    tgt = 1.07 * big @ Rt.T + [4e5, 5e6, 100] + rng.normal(size=(n, 3)) * 0.1
    mask = rng.uniform(size=n) < 0.7
    R1, t1, s1, _ = gsf.umeyama_batched(dev(big), dev(tgt), dev(np.array([0, n]), torch.int64), n, mask=dev(mask.astype(np.uint8), torch.uint8))
    Ro, to, so = fo.umeyama(big[mask], tgt[mask])
    np.testing.assert_allclose(R1[0].cpu().numpy(), Ro, atol=ROT_ATOL); assert abs(float(s1[0].cpu()) - so) < 1e-11
    np.testing.assert_allclose(t1[0].cpu().numpy(), to, rtol=1e-11)
    # fewer than 3 points -> (None, None, None) status
    _, _, _, st2 = gsf.umeyama_batched(dev(big[:2]), dev(tgt[:2]), dev(np.array([0, 2]), torch.int64), 2)
    assert int(st2.cpu()[0]) == 1


@pytest.mark.parametrize("case", ["ransac_outliers", "ransac_collinear", "ransac_hopeless"])
def test_sim3_ransac_matches_seeded_reference(gsf, case):
    """gsf_sim3_ransac_dev with the sample indices of a seeded reference run against that run's
    result (tests/golden/ransac_*.npz from the unmodified reference), then the drop-in function,
    which draws the indices itself from numpy's global RNG."""
    from gps_optimize_slam_b200 import _lib
    g = load_golden(case)
    src, dst, smp = g["src"], g["dst"], g["samples"]
    R, t, s, mask, info, st = gsf.sim3_ransac(dev(src), dev(dst), dev(smp, torch.int32), float(g["thr"]), int(g["min_inliers"]))
    info = info.cpu().numpy(); mask = mask.cpu().numpy().astype(bool)
    if not bool(g["ok"]):
        assert info[2] == 0 and not mask.any() and int(st.cpu()[0]) & _lib.ST_TOO_FEW_POINTS
    else:
        assert info[2] == 1 and info[0] == mask.sum()
        assert not (mask & g["outliers"]).any()                  # no gross outlier survives
        np.testing.assert_allclose(R.cpu().numpy(), g["R"], rtol=0, atol=ROT_ATOL)
        np.testing.assert_allclose(t.cpu().numpy(), g["t"], rtol=1e-11, atol=POS_ATOL)
        assert abs(float(s.cpu()[0]) - float(g["s"])) < 1e-11
    import EKFGPSSLAM as drop
    np.random.seed(int(g["seed"]))
    import contextlib, io
    with contextlib.redirect_stdout(io.StringIO()):
        out = drop.compute_sim3_transform_robust(src, dst, smp.shape[1], float(g["thr"]), len(smp), int(g["min_inliers"]))
    if not bool(g["ok"]):
        assert out == (None, None, None)
    else:
        np.testing.assert_allclose(out[0], g["R"], rtol=0, atol=ROT_ATOL)
        np.testing.assert_allclose(out[1], g["t"], rtol=1e-11, atol=POS_ATOL)
        assert abs(out[2] - float(g["s"])) < 1e-11


def test_ate_kernel(gsf):
    from oracle import fusion_oracle as fo
    cfg = fo.default_config()
    trajs = [t for t in mixed_batch() if len(t["ts"]) >= 64]
    ts, pos, quat, z, off_d, off, max_len = pack(trajs)
    p, q, sim3, st = gsf.fuse_batched(ts, pos, quat, z, off_d, max_len, gsf.params_tensor(cfg))
    stats = gsf.ate_nn_batched(p, z, ts, off_d, max_len, 5.0).cpu().numpy()
    p = p.cpu().numpy()
    for b, tr in enumerate(trajs):
        valid = ~np.isnan(tr["gps"]).any(1)
        ev = fo.evaluation_indices(tr["ts"], valid)
        exp = fo.error_stats(fo.nn_errors(p[off[b]:off[b + 1]], tr["gps"], ev))
        assert int(stats[b, 3]) == len(ev)
        np.testing.assert_allclose(stats[b, :3], exp, rtol=1e-10)


@pytest.mark.parametrize("case", ["pairA", "pairB"])
def test_golden_ate_numbers(gsf, case):
    g = load_golden(case)
    n = len(g["slam_ts"])
    off = dev(np.array([0, n]), torch.int64)
    for row, traj in zip(g["stats"], (g["slam_pos"], g["sim3_pos"], g["ekf_pos"])):
        got = gsf.ate_nn_batched(dev(traj), dev(g["aligned"]), dev(g["slam_ts"]), off, n, 5.0).cpu().numpy()[0]
        np.testing.assert_allclose(got[:3], row, rtol=1e-10)
        assert int(got[3]) == len(g["eval_indices"])


def test_utm_kernels(gsf):
    from oracle import utm_kruger as uk
    rng = np.random.default_rng(2)
    lon = 9.0 + rng.uniform(-3.4, 3.4, 5000); lat = rng.uniform(-79, 83, 5000)
    for south in (False, True):
        e, n = gsf.utm_forward(dev(lon), dev(lat), 32, south)
        eo, no = uk.utm_forward(lon, lat, 32, south)
        np.testing.assert_allclose(e.cpu().numpy(), eo, rtol=0, atol=2e-8)
        np.testing.assert_allclose(n.cpu().numpy(), no, rtol=0, atol=2e-8)
        lo2, la2 = gsf.utm_inverse(e, n, 32, south)
        np.testing.assert_allclose(lo2.cpu().numpy(), lon, rtol=0, atol=1e-12)
        np.testing.assert_allclose(la2.cpu().numpy(), lat, rtol=0, atol=1e-12)
    # far outside the zone (6.9 .. 20 deg from the central meridian): the library-function branch of the forward kernel
    lon_far = 9.0 + rng.choice([-1.0, 1.0], 2000) * rng.uniform(6.8, 20.0, 2000); lat_far = rng.uniform(-79, 83, 2000)
    e, n = gsf.utm_forward(dev(lon_far), dev(lat_far), 32, False)
    eo, no = uk.utm_forward(lon_far, lat_far, 32, False)
    np.testing.assert_allclose(e.cpu().numpy(), eo, rtol=0, atol=2e-8)
    np.testing.assert_allclose(n.cpu().numpy(), no, rtol=0, atol=2e-8)
    for case in ("pairA", "pairB"):
        g = load_golden(case)
        raw = g["gnss_raw"]
        z = gsf.geo_zone(dev(raw[:, 2]), dev(raw[:, 1])).cpu().numpy()
        zone, south = uk.utm_zone_from_means(raw[:, 2], raw[:, 1])
        assert int(z[2]) == zone and bool(z[3]) == south
        e, n = gsf.utm_forward(dev(raw[:, 2]), dev(raw[:, 1]), zone, south)
        np.testing.assert_allclose(np.column_stack((e.cpu().numpy(), n.cpu().numpy())), g["gps_utm"][:, :2], rtol=0, atol=2e-8)


def test_utm_kernels_match_40_digit_fixture(gsf):
    """K1 forward / inverse against tests/golden/utm_mp.npz: the projection's definition evaluated with mpmath at 40
    digits (oracle/utm_mp.py; no coefficient table involved) on a +-3.5 deg x +-84 deg lattice in zones 1 / 32 / 39 / 60,
    both hemispheres.  Tolerance: 3 ulp of the output coordinate (1 ulp = 1.9e-9 m at a 9.3e6 m northing)."""
    d = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "utm_mp.npz"))
    P = d["points"]
    for zone in np.unique(P[:, 2]):
        for south in (0.0, 1.0):
            m = (P[:, 2] == zone) & (P[:, 3] == south)
            if not m.any():
                continue
            e, n = gsf.utm_forward(dev(P[m, 0]), dev(P[m, 1]), int(zone), bool(south))
            np.testing.assert_allclose(e.cpu().numpy(), P[m, 4], rtol=0, atol=6e-9)
            np.testing.assert_allclose(n.cpu().numpy(), P[m, 5], rtol=0, atol=6e-9)
            lon, lat = gsf.utm_inverse(dev(P[m, 4]), dev(P[m, 5]), int(zone), bool(south))
            np.testing.assert_allclose(lon.cpu().numpy(), P[m, 0], rtol=0, atol=1e-12)
            np.testing.assert_allclose(lat.cpu().numpy(), P[m, 1], rtol=0, atol=1e-12)


def test_fused_gnss_ingest_matches_loader_and_oracle(gsf):
    """gsf_gnss_rows_to_utm_dev against the shipped GNSS files as the reference's loader sees them
    (golden gps_utm) and against the oracle on rows with invalid entries / a southern-hemisphere track."""
    from oracle import utm_kruger as uk
    for case in ("pairA", "pairB"):
        g = load_golden(case)
        ts, xyz, zone = gsf.gnss_rows_to_utm(dev(g["gnss_raw"]))
        keep = uk.gnss_validity_mask(g["gnss_raw"][:, 1], g["gnss_raw"][:, 2])
        z = zone.cpu().numpy()
        assert f"{int(z[2])}{'S' if z[3] else 'N'}" == str(g["utm_zone"]) and int(z[4]) == keep.sum()
        np.testing.assert_allclose(xyz.cpu().numpy()[keep], g["gps_utm"], rtol=0, atol=1e-8)
        np.testing.assert_array_equal(ts.cpu().numpy(), g["gnss_raw"][:, 0])
    rng = np.random.default_rng(17)
    n = 100003
    rows = np.column_stack([np.arange(n) * 0.1, -33.9 + rng.normal(0, 0.05, n), 151.2 + rng.normal(0, 0.05, n), 30 + rng.normal(0, 2, n)])
    rows[5, 1] = 0.0; rows[77, 2] = 0.0; rows[500, 1] = 91.0; rows[900, 2] = -181.0
    keep = uk.gnss_validity_mask(rows[:, 1], rows[:, 2])
    zone_o, south_o = uk.utm_zone_from_means(rows[keep, 2], rows[keep, 1])
    e, nn = uk.utm_forward(rows[keep, 2], rows[keep, 1], zone_o, south_o)
    _, xyz, zone = gsf.gnss_rows_to_utm(dev(rows), want_ts=False)
    xyz = xyz.cpu().numpy(); z = zone.cpu().numpy()
    assert int(z[2]) == zone_o and bool(z[3]) == south_o and south_o
    assert np.isnan(xyz[~keep]).all()
    np.testing.assert_allclose(xyz[keep], np.column_stack((e, nn, rows[keep, 3])), rtol=0, atol=1e-8)
    # a track that starts in zone 32 / north and whose overall mean lies in zone 33 / south: the projection runs with the zone
    # of the first rows while it forms the means, and has to be repeated with the zone of the whole track (:127-134)
    n = 20000
    lon = np.where(np.arange(n) < 6000, 11.5, 13.5) + rng.normal(0, 0.01, n)
    lat = np.where(np.arange(n) < 6000, 0.4, -0.4) + rng.normal(0, 0.01, n)
    rows = np.column_stack([np.arange(n) * 0.1, lat, lon, np.full(n, 12.0)])
    zone_o, south_o = uk.utm_zone_from_means(lon, lat)
    assert (zone_o, south_o) == (33, True) and uk.utm_zone_from_means(lon[:4096], lat[:4096]) == (32, False)
    e, nn = uk.utm_forward(lon, lat, zone_o, south_o)
    _, xyz, zone = gsf.gnss_rows_to_utm(dev(rows), want_ts=False)
    z = zone.cpu().numpy()
    assert int(z[2]) == 33 and bool(z[3]) and int(z[4]) == n
    np.testing.assert_allclose(xyz.cpu().numpy(), np.column_stack((e, nn, rows[:, 3])), rtol=0, atol=1e-8)


@pytest.mark.parametrize("case", GOLDEN_CASES)
def test_association_kernel_matches_reference_golden(gsf, case):
    g = load_golden(case)
    one = lambda n: dev(np.array([0, n]), torch.int64)
    a, v = gsf.associate_spline(dev(g["gps_ts"]), dev(g["gps_utm"]), one(len(g["gps_ts"])), dev(g["slam_ts"]), one(len(g["slam_ts"])), 5.0)
    a, v = a.cpu().numpy(), v.cpu().numpy().astype(bool)
    assert np.array_equal(v, g["valid"])
    assert np.isnan(a[~v]).all()
    np.testing.assert_allclose(a[v], g["aligned"][v], rtol=0, atol=POS_ATOL)


def test_association_long_trajectory_local_halo(gsf):
    """gsf_associate_spline_long_dev (local-halo solve, 13-knot chunks in registers) against scipy's interp1d per segment
    (what dynamic_time_alignment calls, EKFGPSSLAM.py:351-380) and against the serial per-trajectory kernel: irregular
    knot spacing, gaps that cut segments of 1, 2, 3, 4, 5, 31, 32, 33, 64, 65 and thousands of knots, stamps exactly on
    knots / segment ends / inside gaps / outside the track; plus the golden cases (271 knots)."""
    from oracle import fusion_oracle as fo
    rng = np.random.default_rng(21)
    lens = [1, 2, 3, 4, 5, 31, 32, 33, 64, 65, 97, 3000, 1, 7000, 4, 20000]
    ts, t0 = [], 100.0
    for m in lens:
        steps = rng.uniform(0.02, 0.4, m) * rng.choice([1.0, 1.0, 5.0], m)
        steps = np.minimum(steps, 4.0)
        seg = t0 + np.cumsum(steps)
        ts.append(seg); t0 = seg[-1] + rng.uniform(5.5, 9.0)
    gt = np.concatenate(ts)
    M = len(gt)
    gy = np.column_stack([455779.0 + 8.0 * (gt - 100.0) + 30 * np.sin(gt / 7.0), 5431368.0 + 50 * np.cos(gt / 11.0) + 3.0 * (gt - 100.0),
                          112.0 + np.sin(gt / 3.0)]) + rng.normal(0, 0.3, (M, 3))
    st = np.concatenate([rng.uniform(gt[0] - 3, gt[-1] + 3, 40000), gt[::7], gt[[0, -1]], [ts[3][0], ts[3][-1], ts[5][-1], ts[11][0]],
                         [ts[1][0] + 1e-7, ts[0][0]]])
    want, want_valid = fo.associate(st, gt, gy, 5.0)
    a, v, status = gsf.associate_spline_long(dev(gt), dev(gy), dev(st), 5.0)
    assert int(status.cpu()[0]) == 0
    a, v = a.cpu().numpy(), v.cpu().numpy().astype(bool)
    np.testing.assert_array_equal(v, want_valid)
    np.testing.assert_allclose(a[v], want[v], rtol=0, atol=POS_ATOL)
    a2, v2 = gsf.associate_spline(dev(gt), dev(gy), dev(np.array([0, M]), torch.int64), dev(st), dev(np.array([0, len(st)]), torch.int64), gap=5.0)
    np.testing.assert_array_equal(v2.cpu().numpy().astype(bool), v)
    np.testing.assert_allclose(a[v], a2.cpu().numpy()[v], rtol=0, atol=2e-8)
    # sorted stamps: narrow brackets, the evaluation searches the knot times staged in shared memory (random order above: the
    # whole-array search); and knot arrays that start 8 bytes off a 16-byte boundary (no TMA staging in the moments kernel)
    order = np.argsort(st, kind="stable")
    a3, v3, _ = gsf.associate_spline_long(dev(gt), dev(gy), dev(st[order]), 5.0)
    np.testing.assert_array_equal(v3.cpu().numpy().astype(bool), v[order])
    np.testing.assert_array_equal(a3.cpu().numpy()[v[order]], a[order][v[order]])
    gt_off, gy_off = dev(np.concatenate([[0.0], gt]))[1:], dev(np.concatenate([np.zeros(3), gy.ravel()]))[3:].view(-1, 3)
    assert gt_off.data_ptr() % 16 == 8 and gy_off.data_ptr() % 16 == 8
    a4, v4, _ = gsf.associate_spline_long(gt_off, gy_off, dev(st[order]), 5.0)
    np.testing.assert_array_equal(a4.cpu().numpy()[v[order]], a3.cpu().numpy()[v[order]])
    np.testing.assert_array_equal(v4.cpu().numpy(), v3.cpu().numpy())
    for case in GOLDEN_CASES:
        g = load_golden(case)
        a, v, _ = gsf.associate_spline_long(dev(g["gps_ts"]), dev(g["gps_utm"]), dev(g["slam_ts"]), 5.0)
        np.testing.assert_array_equal(v.cpu().numpy().astype(bool), g["valid"])
        np.testing.assert_allclose(a.cpu().numpy()[g["valid"]], g["aligned"][g["valid"]], rtol=0, atol=POS_ATOL)


def test_warp_kernel_matches_role_kernel(gsf, monkeypatch):
    """The one-warp-per-trajectory variant of the fast kernel (GSF_FAST_CT=1, kept as a measured alternative) runs the same
    device functions in the same order as the role kernel: bit-identical outputs, ragged lengths and odd offsets included."""
    lens = [271, 33, 1, 2, 288, 97, 270, 3, 64, 255] * 40
    ts, pos, quat, z, off = _ragged_batch(gsf, lens)
    prm = gsf.params_tensor()
    monkeypatch.delenv("GSF_FAST_CT", raising=False)
    a = [o.clone() for o in gsf.fuse_batched(ts, pos, quat, z, off, max(lens), prm)]
    monkeypatch.setenv("GSF_FAST_CT", "1")
    b = gsf.fuse_batched(ts, pos, quat, z, off, max(lens), prm)
    for x, y in zip(a, b):            # bit patterns: the deferred 1- and 2-pose trajectories come out as NaN rows in both
        assert torch.equal(x.view(torch.int64) if x.dtype == torch.float64 else x, y.view(torch.int64) if y.dtype == torch.float64 else y)
    assert int((a[3] == 0).sum()) == 280          # 1, 2 and 3 poses: not enough points for the Sim3 (status set, same in both)


def test_sim3_partial_stats_merge_equals_whole_fit(gsf):
    """A trajectory cut into shards (gsf_sim3_partial_stats_dev per shard, gsf_sim3_from_partial_stats_dev over all) gives the
    Umeyama fit of the whole point set (gsf_sim3_umeyama_batched_dev): masks, an empty shard and a one-point shard included."""
    rng = np.random.default_rng(12)
    n = 50000
    src = np.cumsum(rng.normal(0, 1.0, (n, 3)), axis=0)
    th = 0.7; Rg = np.array([[np.cos(th), -np.sin(th), 0], [np.sin(th), np.cos(th), 0], [0, 0, 1.0]])
    dst = 1.07 * src @ Rg.T + [455000.0, 5431000.0, 110.0] + rng.normal(0, 0.3, (n, 3))
    mask = rng.uniform(size=n) > 0.1
    dst[~mask] = np.nan
    R, t, s, st = gsf.umeyama_batched(dev(src), dev(dst), dev(np.array([0, n]), torch.int64), n, mask=dev(mask, torch.uint8))
    cuts = [0, 17000, 17000, 17001, 40000, n]                       # shards of 17000, 0, 1, 22999, 10000 points
    parts = [gsf.sim3_partial_stats(dev(src[a:b]), dev(dst[a:b]), mask=dev(mask[a:b], torch.uint8)) if b > a
             else torch.zeros(gsf.SIM3_STATS, dtype=torch.float64, device="cuda") for a, b in zip(cuts[:-1], cuts[1:])]
    R2, t2, s2, st2 = gsf.sim3_from_partial_stats(torch.stack(parts))
    assert int(st.cpu()[0]) == 0 and int(st2.cpu()[0]) == 0
    assert float(parts[0][0]) == mask[:17000].sum() and float(parts[2][0]) == float(mask[17000])
    np.testing.assert_allclose(R2.cpu().numpy(), R.cpu().numpy()[0], rtol=0, atol=1e-10)       # (the merge order of the sums differs)
    np.testing.assert_allclose(s2.cpu().numpy(), s.cpu().numpy(), rtol=1e-11, atol=0)
    np.testing.assert_allclose(t2.cpu().numpy(), t.cpu().numpy()[0], rtol=0, atol=1e-6)


def test_fused_call_replays_from_a_cuda_graph(gsf):
    """gsf_fuse_batched_dev inside a stream capture (its memset and both kernels become graph nodes): replays give the
    eager call's bits, also for a batch with deferred (outage) trajectories.  bench.py times config 2 this way."""
    B, n = 600, 271
    ts, pos, quat, z = gsf.synth_generate(B, n, 0.104, 13.0, seed=3, outage_prob=0.3, outage_max_len=15)
    off = gsf.equal_offsets(B, n); prm = gsf.params_tensor()
    want = [o.clone() for o in gsf.fuse_batched(ts, pos, quat, z, off, n, prm)]
    outs = dict(out_pos=torch.empty_like(pos), out_quat=torch.empty_like(quat), sim3_out=torch.empty((B, 16), dtype=torch.float64, device="cuda"),
                status=torch.empty((B,), dtype=torch.int32, device="cuda"))
    gsf.fuse_batched(ts, pos, quat, z, off, n, prm, **outs)                  # warm-up: attributes and occupancy are cached
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        gsf.fuse_batched(ts, pos, quat, z, off, n, prm, **outs)
    for _ in range(3):
        for t in outs.values():
            t.zero_()
        g.replay()
        torch.cuda.synchronize()
        got = (outs["out_pos"], outs["out_quat"], outs["sim3_out"], outs["status"])
        for x, y in zip(want, got):
            assert torch.equal(x.view(torch.int64) if x.dtype == torch.float64 else x, y.view(torch.int64) if y.dtype == torch.float64 else y)
    assert 0 < int((want[3] == 0).sum()) <= B


def _ragged_batch(gsf, lens):
    """Device-generated equal-length batch cut to ragged lengths (offsets into the packed arrays)."""
    n = max(lens)
    ts, pos, quat, z = gsf.synth_generate(len(lens), n, 0.104, 13.0, seed=11)
    keep = torch.zeros((len(lens), n), dtype=torch.bool, device=ts.device)
    for k, m in enumerate(lens):
        keep[k, :m] = True
    flat = keep.reshape(-1)
    off = torch.tensor(np.concatenate([[0], np.cumsum(lens)]), dtype=torch.int64, device=ts.device)
    return ts.reshape(-1)[flat].contiguous(), pos.reshape(-1, 3)[flat].contiguous(), quat.reshape(-1, 4)[flat].contiguous(), z.reshape(-1, 3)[flat].contiguous(), off


def test_association_long_gaps_at_chunk_and_block_edges(gsf):
    """Segment ends placed on and next to the boundaries of the moments kernel's work split (13-knot chunks, 1664-knot blocks; the 15 / 1920 of an earlier layout stay in the list,
    22-knot window margins, 20-knot halos) and next to the ends of the track, sorted stamps (staged evaluation) that include
    every knot: against scipy per segment (oracle)."""
    from oracle import fusion_oracle as fo
    rng = np.random.default_rng(33)
    M = 6000
    edge_sets = [
        [1663, 1664, 1665], [1651, 1677], [1642, 1643, 1644, 1685, 1686], [12, 13, 14, 25, 26, 27], [3327, 3328, 3329, 3341], [1664 - 20, 1664 + 19],
        [1664 - 21, 1664 + 20, 1664 + 21], list(range(1640, 1700, 4)),
        [1919, 1920, 1921], [1905, 1935], [1898, 1899, 1900, 1941, 1942], [14, 15, 16, 29, 30, 31], [3839, 3840, 3841, 3855, 3860],
        [1, 2, 3], [4, 5996], [5995, 5996, 5997, 5998], [1920 - 20, 1920 + 19], [1920 - 21, 1920 + 20, 1920 + 21], [3, 7, 11, 15, 19, 23, 27],
        list(range(1900, 1960, 4)), [],
    ]
    for edges in edge_sets:
        steps = rng.uniform(0.05, 0.3, M)
        steps[np.array(edges, dtype=int)] = rng.uniform(5.5, 8.0, len(edges))       # a gap BEFORE knot k for every k in edges
        steps[0] = 0.1
        gt = 50.0 + np.cumsum(steps)
        gy = np.column_stack([4.5e5 + 9.0 * gt + 20 * np.sin(gt / 5.0), 5.4e6 + 40 * np.cos(gt / 9.0), 100.0 + np.sin(gt / 2.0)]) + rng.normal(0, 0.3, (M, 3))
        st = np.sort(np.concatenate([rng.uniform(gt[0] - 1, gt[-1] + 1, 9000), gt]))
        want, want_valid = fo.associate(st, gt, gy, 5.0)
        a, v, status = gsf.associate_spline_long(dev(gt), dev(gy), dev(st), 5.0)
        assert int(status.cpu()[0]) == 0
        a, v = a.cpu().numpy(), v.cpu().numpy().astype(bool)
        np.testing.assert_array_equal(v, want_valid, err_msg=str(edges))
        np.testing.assert_allclose(a[v], want[v], rtol=0, atol=POS_ATOL, err_msg=str(edges))


def test_association_long_degenerate_sizes(gsf):
    """Tracks of 1 .. 6 knots, stamp sets of 1 / 3 / 2049 / 4100 entries (tile edges of the evaluation), NaN stamps, and the
    status flag for stamps that do not increase inside a segment (:356-359)."""
    from oracle import fusion_oracle as fo
    rng = np.random.default_rng(5)
    for M in (1, 2, 3, 4, 5, 6):
        gt = 10.0 + np.cumsum(rng.uniform(0.1, 0.5, M))
        gy = rng.normal(0, 5.0, (M, 3)) + [4e5, 5e6, 100.0]
        for N in (1, 3, 2049, 4100):
            st = np.sort(rng.uniform(gt[0] - 0.5, gt[-1] + 0.5, N))
            st[N // 2] = gt[M // 2]
            want, want_valid = fo.associate(st, gt, gy, 5.0)
            a, v, status = gsf.associate_spline_long(dev(gt), dev(gy), dev(st), 5.0)
            assert int(status.cpu()[0]) == 0
            a, v = a.cpu().numpy(), v.cpu().numpy().astype(bool)
            np.testing.assert_array_equal(v, want_valid, err_msg=f"M={M} N={N}")
            np.testing.assert_allclose(a[v], want[v], rtol=0, atol=POS_ATOL, err_msg=f"M={M} N={N}")
            assert np.isnan(a[~v]).all()
    gt = 10.0 + np.cumsum(rng.uniform(0.1, 0.5, 5000)); gy = rng.normal(0, 5.0, (5000, 3))
    st = np.sort(rng.uniform(gt[0], gt[-1], 3000)); st[[0, 7, 2999]] = np.nan
    a, v, status = gsf.associate_spline_long(dev(gt), dev(gy), dev(st), 5.0)
    v = v.cpu().numpy().astype(bool)
    assert not v[[0, 7, 2999]].any() and v.sum() == 2997 and int(status.cpu()[0]) == 0
    gt_bad = gt.copy(); gt_bad[2500] = gt_bad[2499]                       # a repeated stamp inside a segment
    _, _, status = gsf.associate_spline_long(dev(gt_bad), dev(gy), dev(st), 5.0)
    assert int(status.cpu()[0]) == 1


def test_association_short_segments(gsf):
    """2-3 knot segments are linear, 1-knot segments are skipped (EKFGPSSLAM.py:361-362)."""
    from oracle import fusion_oracle as fo
    gt = np.array([0.0, 1.0, 2.0, 10.0, 11.0, 20.0, 30.0, 31.0, 32.0, 33.0, 34.5])
    gy = np.cumsum(np.random.default_rng(0).normal(size=(len(gt), 3)), axis=0) + [4e5, 5e6, 100]
    st = np.linspace(-1, 36, 400)
    ao, vo = fo.associate(st, gt, gy)
    one = lambda n: dev(np.array([0, n]), torch.int64)
    a, v = gsf.associate_spline(dev(gt), dev(gy), one(len(gt)), dev(st), one(len(st)), 5.0)
    assert np.array_equal(v.cpu().numpy().astype(bool), vo)
    np.testing.assert_allclose(a.cpu().numpy()[vo], ao[vo], rtol=0, atol=POS_ATOL)


def test_device_generator_slice_matches_oracle(gsf):
    """Config-3-style data generated on the device, a slice checked against the oracle; the
    generator is counter-based so a shard regenerates identically."""
    from oracle import fusion_oracle as fo
    cfg = fo.default_config()
    B, n = 256, 1000
    ts, pos, quat, z = gsf.synth_generate(B, n, 0.1, 10.0, seed=1234, outage_prob=0.3, outage_max_len=120)
    off = gsf.equal_offsets(B, n)
    p, q, sim3, st = gsf.fuse_batched(ts, pos, quat, z, off, n, gsf.params_tensor(cfg))
    ts2, pos2, quat2, z2 = gsf.synth_generate(64, n, 0.1, 10.0, seed=1234, first_traj=128, outage_prob=0.3, outage_max_len=120)
    assert torch.equal(pos2, pos[128 * n:192 * n]) and torch.equal(quat2, quat[128 * n:192 * n])
    h = [x.cpu().numpy() for x in (ts, pos, quat, z, p, q, sim3, st)]
    checked = 0
    for b in list(range(0, 12)) + [100, 255]:
        sl = slice(b * n, (b + 1) * n)
        tr = dict(ts=h[0][sl], pos=h[1][sl], quat=h[2][sl], gps=h[3][sl])
        o = oracle_pipeline(tr, cfg)
        assert h[7][b] == 0
        np.testing.assert_allclose(h[4][sl], o["pos"], rtol=0, atol=POS_ATOL)
        np.testing.assert_allclose(h[5][sl], o["quat"], rtol=0, atol=ROT_ATOL)
        np.testing.assert_allclose(h[6][b, :9].reshape(3, 3), o["R"], atol=ROT_ATOL)
        checked += 1
    assert checked == 14 and np.isnan(h[3]).any()


def test_host_buffer_entry_matches_device_entry(gsf):
    from gps_optimize_slam_b200.config import pack_fuse_params
    trajs = mixed_batch() * 3
    ts, pos, quat, z, off_d, off, max_len = pack(trajs)
    p, q, sim3, st = gsf.fuse_batched(ts, pos, quat, z, off_d, max_len, gsf.params_tensor())
    cat = lambda k: np.ascontiguousarray(np.concatenate([t[k] for t in trajs]))
    hp, hq, hs, hst = gsf.fuse_batched_host(cat("ts"), cat("pos"), cat("quat"), cat("gps"), off, max_len, pack_fuse_params())
    assert np.array_equal(hp, p.cpu().numpy()) and np.array_equal(hq, q.cpu().numpy())
    assert np.array_equal(hst, st.cpu().numpy()) and np.array_equal(hs[:, :15], sim3.cpu().numpy()[:, :15])


def test_per_trajectory_noise_parameters(gsf):
    from gps_optimize_slam_b200 import synth
    from oracle import fusion_oracle as fo
    tr = synth.make_trajectory(31, n=200, outages=[(60, 90)])
    sets = [([0.1] * 3 + [0.01] * 4, [q, q, 3 * q] + [0.01] * 4, [r] * 3) for q in (0.01, 0.1, 1.0) for r in (0.05, 0.2, 2.0)]
    ts, pos, quat, z, off_d, off, max_len = pack([tr] * len(sets))
    prm = gsf.params_tensor(per_traj=sets)
    p, q, _, st = gsf.fuse_batched(ts, pos, quat, z, off_d, max_len, prm, params_per_traj=True)
    p = p.cpu().numpy()
    for b, (p0, qq, rr) in enumerate(sets):
        cfg = fo.default_config()
        cfg["ekf"].update(initial_cov_diag=p0, process_noise_diag=qq, meas_noise_diag=rr)
        np.testing.assert_allclose(p[off[b]:off[b + 1]], oracle_pipeline(tr, cfg)["pos"], rtol=0, atol=POS_ATOL)


def test_dropin_ekf_object_rts_nlerp_sharp_turn(gsf):
    """The per-call surface of the drop-in (ExtendedKalmanFilter.process_step / _predict / _update,
    rts_smoother_segment, quaternion_nlerp, is_sharp_turn_in_segment: gsf_ekf_step_dev, gsf_rts_segment_dev,
    gsf_quat_nlerp_dev, gsf_sharp_turn_dev) against the oracle's dense 7x7 restatement, step by step over a trajectory
    with an outage, a recovery and blended updates."""
    import contextlib, io
    import EKFGPSSLAM as drop
    from gps_optimize_slam_b200 import synth
    from oracle import fusion_oracle as fo
    tr = synth.make_trajectory(91, n=60, outages=[(20, 33)])
    cfg = fo.default_config()["ekf"]
    p0, q0 = tr["gps"][0], tr["quat"][0]
    mine = drop.ExtendedKalmanFilter(p0, q0, cfg)
    ref = fo.DenseEKF(p0, q0, cfg)
    hist = {k: [] for k in ("xf", "Pf", "xp", "Pp")}
    for i in range(1, 60):
        motion = fo.relative_pose(tr["pos"][i - 1], tr["quat"][i - 1], tr["pos"][i], tr["quat"][i])
        z = tr["gps"][i]
        avail = not np.isnan(z).any()
        steps = 3 if 33 <= i < 40 else 0                       # blended updates right after the recovery
        a = mine.process_step(motion, z if avail else None, avail, tr["ts"][i] - tr["ts"][i - 1], steps)
        b = ref.step(motion, z if avail else None, avail, tr["ts"][i] - tr["ts"][i - 1], steps)
        for got, want, atol in zip(a, b, (POS_ATOL, 1e-12, POS_ATOL, 1e-12)):
            np.testing.assert_allclose(got, want, rtol=0, atol=atol)
        assert abs(mine.gnss_update_weight - ref.weight) < 1e-15
        for k, v in zip(("xf", "Pf", "xp", "Pp"), a):
            hist[k].append(np.array(v))
    xp, Pp = mine._predict(ref.x, ref.P, motion, 0.1)
    xo, Po = ref.predict(motion, 0.1)
    np.testing.assert_allclose(xp, xo, rtol=0, atol=POS_ATOL); np.testing.assert_allclose(Pp, Po, rtol=0, atol=1e-12)
    xu, Pu = mine._update(xo, Po, tr["gps"][5])
    xuo, Puo = ref.update(xo, Po, tr["gps"][5])
    np.testing.assert_allclose(xu, xuo, rtol=0, atol=POS_ATOL); np.testing.assert_allclose(Pu, Puo, rtol=0, atol=1e-12)
    assert mine._update(xo, Po, np.array([np.nan, 0.0, 0.0])) == (None, None)
    # RTS over the outage segment of the recorded history (dense covariances with off-diagonal terms added)
    sl = slice(18, 34)
    rng = np.random.default_rng(2)
    def spd(P):
        A = rng.normal(size=(7, 7)) * 1e-3
        return P + A @ A.T
    Pf = [spd(P) for P in hist["Pf"][sl]]; Pp_ = [spd(P) for P in hist["Pp"][sl]]
    xs, Ps = drop.rts_smoother_segment(hist["xf"][sl], Pf, hist["xp"][sl], Pp_)
    want = fo.rts_segment(hist["xf"][sl], Pf, hist["xp"][sl], Pp_)
    for got, w_ in zip(xs, want):
        np.testing.assert_allclose(got, w_, rtol=0, atol=POS_ATOL)
    assert len(Ps) == len(xs) and all(np.allclose(P, P.T) for P in Ps)
    # nlerp (sign flip, clipping, degenerate antipodal midpoint) and the sharp-turn gate
    qa = np.array([0.1, -0.2, 0.3, 0.9]); qa /= np.linalg.norm(qa)
    qb = -np.array([0.12, -0.18, 0.33, 0.88]); qb /= np.linalg.norm(qb)
    for w in (-0.3, 0.0, 0.25, 0.5, 1.0, 1.7):
        np.testing.assert_allclose(drop.quaternion_nlerp(qa, qb, w), fo.nlerp(qa, qb, w), rtol=0, atol=1e-15)
    for w in (0.5, 0.2, 0.9):
        np.testing.assert_allclose(drop.quaternion_nlerp(qa, -qa * 1.0, w), fo.nlerp(qa, -qa * 1.0, w), rtol=0, atol=1e-15)
    with contextlib.redirect_stdout(io.StringIO()):
        for case in (synth.make_trajectory(92, n=40), synth.make_trajectory(93, n=40, sharp_turn_at=20)):
            for a0, a1 in ((0, 40), (15, 26), (3, 4), (7, 8)):
                got = drop.is_sharp_turn_in_segment(list(case["quat"][a0:a1]), list(case["ts"][a0:a1]), np.deg2rad(45.0))
                assert got == fo.sharp_turn(case["quat"][a0:a1], case["ts"][a0:a1], np.deg2rad(45.0))
        bad = case["quat"][10:20].copy(); bad[4] = 0.0
        assert drop.is_sharp_turn_in_segment(list(bad), list(case["ts"][10:20]), np.deg2rad(45.0)) is True
    with pytest.raises(ValueError):
        mine.process_step((np.zeros(3), np.zeros(4)), None, False, 0.1)


def test_fast_kernel_per_trajectory_parameters_all_axes_distinct(gsf):
    """All-valid trajectories (fast warp-specialised kernel) with one parameter record per trajectory, x / y / z
    noise all different (three-axis covariance scan), ragged lengths with odd offsets; also checks that the fast and
    the general kernel (GSF_FUSE_IMPL=general) agree."""
    import os
    from gps_optimize_slam_b200 import synth
    from oracle import fusion_oracle as fo
    rng = np.random.default_rng(77)
    lens = [271, 1000, 333, 64, 999, 1001, 577]
    trajs = [synth.make_trajectory(200 + k, n=n, dt=0.1, speed=10.0) for k, n in enumerate(lens)]
    sets = []
    for _ in lens:
        qv = 10 ** rng.uniform(-2, 0.5, 3); rv = 10 ** rng.uniform(-1.5, 0.5, 3); pv = 10 ** rng.uniform(-2, 0, 3)
        sets.append((list(pv) + [0.01] * 4, list(qv) + [0.01] * 4, list(rv)))
    sets[1] = ([0.1] * 3 + [0.01] * 4, [0.1, 0.1, 0.7] + [0.01] * 4, [0.2] * 3)        # shipped CONFIG: x = y
    ts, pos, quat, z, off_d, off, max_len = pack(trajs)
    prm = gsf.params_tensor(per_traj=sets)
    p, q, sim3, st = gsf.fuse_batched(ts, pos, quat, z, off_d, max_len, prm, params_per_traj=True)
    assert (st.cpu().numpy() == 0).all()
    os.environ["GSF_FUSE_IMPL"] = "general"
    try:
        pg, qg, _, stg = gsf.fuse_batched(ts, pos, quat, z, off_d, max_len, prm, params_per_traj=True)
    finally:
        del os.environ["GSF_FUSE_IMPL"]
    p, q, pg, qg = [x.cpu().numpy() for x in (p, q, pg, qg)]
    np.testing.assert_allclose(p, pg, rtol=0, atol=POS_ATOL); np.testing.assert_allclose(q, qg, rtol=0, atol=ROT_ATOL)
    for b, (p0, qq, rr) in enumerate(sets):
        cfg = fo.default_config()
        cfg["ekf"].update(initial_cov_diag=p0, process_noise_diag=qq, meas_noise_diag=rr)
        o = oracle_pipeline(trajs[b], cfg)
        np.testing.assert_allclose(p[off[b]:off[b + 1]], o["pos"], rtol=0, atol=POS_ATOL)
        np.testing.assert_allclose(q[off[b]:off[b + 1]], o["quat"], rtol=0, atol=ROT_ATOL)


def test_fast_kernel_deferral_conditions_and_work_queue(gsf):
    """Inputs the warp-specialised kernel must hand to the general kernel (ST_DEFERRED internally, never visible to
    the caller), mixed with trajectories it handles itself and with empty ones the work queue has to skip:
    repeated / decreasing timestamps (the reference clamps dt to 1e-6 s, :863), a measurement variance so small that
    the projective covariance recursion leaves its range, and a normal trajectory with extreme but representable
    noise ratios.  Many more trajectories than thread blocks per SM slot, so the dynamic queue wraps its ring."""
    import os
    from gps_optimize_slam_b200 import synth
    from oracle import fusion_oracle as fo
    lens = [271, 300, 0, 271, 64, 1000, 0, 333, 271, 500] + [97] * 30
    trajs = [synth.make_trajectory(500 + k, n=max(n, 1), dt=0.1, speed=10.0) for k, n in enumerate(lens)]
    for k, n in enumerate(lens):
        if n == 0:
            for key in ("ts", "pos", "quat", "gps"):
                trajs[k][key] = trajs[k][key][:0]
    trajs[1]["ts"][100] = trajs[1]["ts"][99]                      # dt = 0      -> clamp in the reference
    trajs[3]["ts"][50] = trajs[3]["ts"][49] - 0.03                # dt < 0      -> clamp in the reference
    trajs[7]["ts"][10] = trajs[7]["ts"][9] + 5e-7                 # 0 < dt < 1e-6
    base = ([0.1] * 3 + [0.01] * 4, [0.1, 0.1, 0.7] + [0.01] * 4, [0.2] * 3)
    sets = [base] * len(lens)
    sets[4] = ([0.1] * 3 + [0.01] * 4, [50.0, 50.0, 80.0] + [0.01] * 4, [1e-14] * 3)      # q dt / r ~ 5e14 per step
    sets[8] = ([1e-3] * 3 + [0.01] * 4, [10.0, 10.0, 3.0] + [0.01] * 4, [1e-4, 1e-4, 2e-4])  # large but in range
    ts, pos, quat, z, off_d, off, max_len = pack(trajs)
    prm = gsf.params_tensor(per_traj=sets)
    p, q, sim3, st = gsf.fuse_batched(ts, pos, quat, z, off_d, max_len, prm, params_per_traj=True)
    os.environ["GSF_FUSE_IMPL"] = "general"
    try:
        pg, qg, _, stg = gsf.fuse_batched(ts, pos, quat, z, off_d, max_len, prm, params_per_traj=True)
    finally:
        del os.environ["GSF_FUSE_IMPL"]
    st, stg = st.cpu().numpy(), stg.cpu().numpy()
    np.testing.assert_array_equal(st, stg)
    assert (st[[k for k, n in enumerate(lens) if n > 0]] == 0).all(), st
    p, q, pg, qg = [x.cpu().numpy() for x in (p, q, pg, qg)]
    np.testing.assert_allclose(p, pg, rtol=0, atol=POS_ATOL); np.testing.assert_allclose(q, qg, rtol=0, atol=ROT_ATOL)
    for b in (0, 1, 3, 4, 5, 7, 8, 12):
        cfg = fo.default_config()
        cfg["ekf"].update(initial_cov_diag=sets[b][0], process_noise_diag=sets[b][1], meas_noise_diag=sets[b][2])
        o = oracle_pipeline(trajs[b], cfg)
        np.testing.assert_allclose(p[off[b]:off[b + 1]], o["pos"], rtol=0, atol=POS_ATOL)
        np.testing.assert_allclose(q[off[b]:off[b + 1]], o["quat"], rtol=0, atol=ROT_ATOL)


@pytest.mark.parametrize("n,dt", [(300, 0.104), (2200, 0.104)])
def test_hypothesis_grid_matches_oracle(gsf, n, dt):
    """gsf_ekf_hypothesis_grid_dev (one trajectory x H noise sets -> ATE statistics) against the oracle run
    hypothesis by hypothesis; n = 2200 spans 229 s, so the 180 s Sim3 window (:985-998) is exercised, and a
    closed loop makes the pruned nearest-neighbour search meet candidates from another lap."""
    from gps_optimize_slam_b200 import synth
    from gps_optimize_slam_b200.config import pack_fuse_params
    from oracle import fusion_oracle as fo
    tr = synth.make_trajectory(41, n=n, dt=dt, speed=9.0)
    if n > 1000:                                   # second lap over the first one: NN candidates from another time
        half = n // 2
        tr["gps"][half:] = tr["gps"][:n - half] + np.random.default_rng(3).normal(0, 0.2, (n - half, 3))
    rng = np.random.default_rng(9)
    sets = [([0.1] * 3 + [0.01] * 4, [0.1, 0.1, 0.7] + [0.01] * 4, [0.2] * 3)]
    for _ in range(40 if n <= 1000 else 9):
        qv = 10 ** rng.uniform(-3, 1, 3); rv = 10 ** rng.uniform(-2, 1, 3); pv = 10 ** rng.uniform(-2, 0, 3)
        sets.append((list(pv) + [0.01] * 4, list(qv) + [0.01] * 4, list(rv)))
    blob = np.concatenate([pack_fuse_params(None, p0=a, q=b, r=c) for (a, b, c) in sets])
    stats, sim3, st = gsf.hypothesis_grid(dev(tr["ts"]), dev(tr["pos"]), dev(tr["quat"]), dev(tr["gps"]), dev(blob, torch.uint8))
    assert int(st.cpu()[0]) == 0
    stats = stats.cpu().numpy()
    valid = np.ones(n, dtype=bool)
    ev = fo.evaluation_indices(tr["ts"], valid)
    for h, (p0, qq, rr) in enumerate(sets):
        cfg = fo.default_config()
        cfg["ekf"].update(initial_cov_diag=p0, process_noise_diag=qq, meas_noise_diag=rr)
        o = oracle_pipeline(tr, cfg)
        want = fo.error_stats(fo.nn_errors(o["pos"], tr["gps"], ev))
        # errors are differences of UTM-scale coordinates (one ulp = 9.3e-10 m): same absolute budget as positions
        np.testing.assert_allclose(stats[h, :3], want, rtol=0, atol=POS_ATOL)
        assert stats[h, 3] == len(ev)
    np.testing.assert_allclose(sim3.cpu().numpy()[:9].reshape(3, 3), o["R"], atol=ROT_ATOL)
    # a pose without GNSS is refused (outage / RTS needs per-hypothesis history)
    bad = tr["gps"].copy(); bad[50] = np.nan
    stats2, _, st2 = gsf.hypothesis_grid(dev(tr["ts"]), dev(tr["pos"]), dev(tr["quat"]), dev(bad), dev(blob, torch.uint8))
    from gps_optimize_slam_b200 import _lib
    assert int(st2.cpu()[0]) & _lib.ST_GRID_NEEDS_ALL_VALID and np.isnan(stats2.cpu().numpy()[:, :3]).all()


@pytest.mark.parametrize("n,k", [(300, (5, 4, 6)), (4541, (3, 3, 4))])
def test_noise_grid_matches_per_hypothesis_kernel_and_oracle(gsf, n, k):
    """gsf_ekf_noise_grid_dev (product grid, factored into scalar x/y and z tracks + a combine kernel) against
    gsf_ekf_hypothesis_grid_dev on the same hypotheses (same per-step arithmetic, tracks parallel in time: equal to
    rounding) and against the oracle (EKFGPSSLAM.py:679-772, :1021-1033); a shard [h_first, h_first + h_count) that
    starts and ends inside a q_xy slab equals the same rows of the full run bit for bit."""
    from gps_optimize_slam_b200 import synth
    from gps_optimize_slam_b200.config import pack_fuse_params, pack_noise_grid
    from oracle import fusion_oracle as fo
    tr = synth.make_loop_trajectory(7, n=n) if n > 1000 else synth.make_trajectory(43, n=n, dt=0.104, speed=9.0)
    Kq, Kz, Kr = k
    qxy = np.logspace(-3, 1, Kq); qz = np.logspace(-2.5, 0.5, Kz); rr = np.logspace(-2, 1, Kr)
    grid = np.stack(np.meshgrid(qxy, qz, rr, indexing="ij"), axis=-1).reshape(-1, 3)
    H = len(grid)
    args = [dev(tr[key]) for key in ("ts", "pos", "quat", "gps")]
    base = dev(pack_fuse_params(), torch.uint8)
    stats, sim3, st = gsf.noise_grid(*args, base, dev(qxy), dev(qz), dev(rr))
    assert int(st.cpu()[0]) == 0
    ref, sim3b, stb = gsf.hypothesis_grid(*args, dev(pack_noise_grid(grid), torch.uint8))
    stats, ref = stats.cpu().numpy(), ref.cpu().numpy()
    # same per-step arithmetic; the tracks are computed parallel in time (chunk-start values from Moebius / affine scans), so they
    # agree with the per-hypothesis kernel's serial recursion to a few 1e-16 relative of the UTM-scale coordinates
    np.testing.assert_array_equal(stats[:, 3], ref[:, 3])
    np.testing.assert_allclose(stats[:, :3], ref[:, :3], rtol=0, atol=2e-8)
    np.testing.assert_array_equal(sim3.cpu().numpy(), sim3b.cpu().numpy())
    # shard inside the grid
    h0, hc = Kz * Kr + 3, H - 2 * Kz * Kr - 5 if Kq > 3 else H - Kz * Kr - 7
    part, _, _ = gsf.noise_grid(*args, base, dev(qxy), dev(qz), dev(rr), h_first=h0, h_count=hc)
    np.testing.assert_array_equal(part.cpu().numpy(), stats[h0:h0 + hc])
    one, _, _ = gsf.noise_grid(*args, base, dev(qxy), dev(qz), dev(rr), h_first=5, h_count=2)   # inside one (q_xy, q_z) row
    np.testing.assert_array_equal(one.cpu().numpy(), stats[5:7])
    # oracle, a few hypotheses
    valid = np.ones(n, dtype=bool)
    ev = fo.evaluation_indices(tr["ts"], valid)
    for h in ([0, H // 2, H - 1] if n > 1000 else range(0, H, 7)):
        cfg = fo.default_config()
        cfg["ekf"].update(process_noise_diag=[grid[h, 0], grid[h, 0], grid[h, 1]] + [0.01] * 4, meas_noise_diag=[grid[h, 2]] * 3)
        o = oracle_pipeline(tr, cfg)
        want = fo.error_stats(fo.nn_errors(o["pos"], tr["gps"], ev))
        np.testing.assert_allclose(stats[h, :3], want, rtol=0, atol=POS_ATOL)
        assert stats[h, 3] == len(ev)
    bad = tr["gps"].copy(); bad[50] = np.nan
    s2, _, st2 = gsf.noise_grid(args[0], args[1], args[2], dev(bad), base, dev(qxy), dev(qz), dev(rr))
    from gps_optimize_slam_b200 import _lib
    assert int(st2.cpu()[0]) & _lib.ST_GRID_NEEDS_ALL_VALID and np.isnan(s2.cpu().numpy()[:, :3]).all()


def test_bit_reproducible_and_shard_invariant(gsf):
    """Same inputs -> same bits; an N-way shard (separate launches over trajectory ranges)
    equals the unsharded run bit for bit (multi-GPU without a cluster)."""
    B, n = 512, 271
    ts, pos, quat, z = gsf.synth_generate(B, n, 0.104, 13.0, seed=9, outage_prob=0.2, outage_max_len=80)
    off = gsf.equal_offsets(B, n)
    prm = gsf.params_tensor()
    p1, q1, s1, _ = gsf.fuse_batched(ts, pos, quat, z, off, n, prm)
    p2, q2, s2, _ = gsf.fuse_batched(ts, pos, quat, z, off, n, prm)
    assert torch.equal(p1, p2) and torch.equal(q1, q2) and torch.equal(s1[:, :15], s2[:, :15])
    for k in range(4):
        a, b = k * B // 4, (k + 1) * B // 4
        sl = slice(a * n, b * n)
        ps, qs, ss, _ = gsf.fuse_batched(ts[sl], pos[sl], quat[sl], z[sl], gsf.equal_offsets(b - a, n), n, prm)
        assert torch.equal(ps, p1[sl]) and torch.equal(qs, q1[sl]) and torch.equal(ss[:, :15], s1[a:b, :15])


def test_full_size_properties(gsf):
    """Config-2 size (4096 x 271) and a config-3 slab (8192 x 1000): size-independent
    properties -- Sim3 recovers a rotation (R R^T = I, det +1), fused track hugs GNSS better
    than the Sim3-aligned one, strict and scan formulations agree."""
    prm = gsf.params_tensor()
    for (B, n, dt, v) in ((4096, 271, 0.104, 13.0), (8192, 1000, 0.1, 10.0)):
        ts, pos, quat, z = gsf.synth_generate(B, n, dt, v, seed=77)
        off = gsf.equal_offsets(B, n)
        p, q, sim3, st = gsf.fuse_batched(ts, pos, quat, z, off, n, prm)
        assert int(st.abs().sum().cpu()) == 0
        R = sim3[:, :9].reshape(B, 3, 3)
        eye = torch.eye(3, dtype=torch.float64, device="cuda")
        assert float((R @ R.transpose(1, 2) - eye).abs().max().cpu()) < 1e-12
        assert float((torch.linalg.det(R) - 1).abs().max().cpu()) < 1e-12
        assert float((q.norm(dim=1) - 1).abs().max().cpu()) < 1e-12
        s = sim3[:, 12]
        assert float(s.min().cpu()) > 0.85 and float(s.max().cpu()) < 1.15
        sp, sq, _ = gsf.sim3_apply_batched(pos, quat, off, n, R.contiguous(), sim3[:, 9:12].contiguous(), s.contiguous())
        err_f = (p - z).norm(dim=1).reshape(B, n)[:, 50:].mean()
        err_s = (sp - z).norm(dim=1).reshape(B, n)[:, 50:].mean()
        assert float(err_f.cpu()) < float(err_s.cpu())
        ps, qs, _ = gsf.ekf_strict_batched(ts, pos, quat, z, off, prm, sp.reshape(B, n, 3)[:, 0].contiguous(), sq.reshape(B, n, 4)[:, 0].contiguous())
        assert float((ps - p).abs().max().cpu()) < POS_ATOL and float((qs - q).abs().max().cpu()) < ROT_ATOL


def test_fuse_trajectory_falls_back_to_ransac_on_outliers(gsf):
    """The single-launch path fits all selected points; with gross GNSS outliers in the Sim3 window the reference's
    compute_sim3_transform_robust (EKFGPSSLAM.py:389-426) refits on an inlier subset.  fuse_trajectory must notice
    (status bit 16) and take the explicit RANSAC route, so that it returns what main_process / the reference return:
    compared with the oracle's RANSAC (same numpy seed -> same sample indices, :408) + EKF seeded with its pose 0."""
    import EKFGPSSLAM as E
    from gps_optimize_slam_b200 import _lib, synth
    from oracle import fusion_oracle as fo
    tr = synth.make_trajectory(77, n=271)
    clean = E.fuse_trajectory({"timestamps": tr["ts"], "positions": tr["pos"], "quaternions": tr["quat"]}, tr["gps"])
    assert clean["status"] == 0 and clean["sim3_path"].startswith("fused")
    z = tr["gps"].copy()
    for i in (20, 90, 150, 151, 230):
        z[i] += np.array([35.0, -28.0, 12.0])
    slam = {"timestamps": tr["ts"], "positions": tr["pos"], "quaternions": tr["quat"]}
    np.random.seed(11)
    out = E.fuse_trajectory(slam, z)
    assert out["status"] & _lib.ST_RANSAC_OUTLIERS and out["sim3_path"].startswith("RANSAC")
    cfg = fo.default_config(); rc = cfg["sim3_ransac"]
    valid = np.ones(271, dtype=bool)
    sel = fo.sim3_point_selection(tr["ts"], valid)
    np.random.seed(11)
    R, t, s = fo.sim3_ransac(tr["pos"][sel], z[sel], rc["min_samples"], rc["residual_threshold"], rc["max_trials"], rc["min_inliers_needed"])
    sp, sq = fo.sim3_apply(tr["pos"], tr["quat"], R, t, s)
    fp, fq = fo.ekf_fuse(tr["ts"], tr["pos"], tr["quat"], z, valid, sp[0], sq[0], cfg)
    np.testing.assert_allclose(out["R"], R, rtol=0, atol=ROT_ATOL)
    assert abs(out["s"] - s) < 1e-11
    np.testing.assert_allclose(out["pos"], fp, rtol=0, atol=POS_ATOL)
    np.testing.assert_allclose(out["quat"], fq, rtol=0, atol=ROT_ATOL)
    # the all-points fit is visibly different here (that is what the flag protects against)
    allpts = fo.umeyama(tr["pos"][sel], z[sel])
    assert np.abs(allpts[1] - t).max() > 1e-3


@pytest.mark.parametrize("case", ["gpsfilter_sliding", "gpsfilter_sparse", "gpsfilter_global", "gpsfilter_clean"])
def test_gps_prefilter_matches_seeded_reference(gsf, case):
    """N3: the drop-in's filter_gps_outliers_ransac (per-window, per-axis polynomial RANSAC on the device,
    gsf_poly_ransac_dev) against the UNMODIFIED reference's (EKFGPSSLAM.py:136-247, sklearn RANSACRegressor) run with the
    same numpy seed (tests/golden/gpsfilter_*.npz, oracle/make_golden_gpsfilter.py): the same points survive and numpy's
    global RNG ends in the same state, i.e. every fit stopped after the same number of trials."""
    import EKFGPSSLAM as E
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", case + ".npz"))
    cfg = dict(E.CONFIG["gps_filtering_ransac"])
    cfg["enabled"] = True
    cfg["use_sliding_window"] = bool(g["sliding"])
    np.random.seed(int(g["seed"]))
    tf, pf = E.filter_gps_outliers_ransac(g["t"], g["pos"], cfg)
    np.testing.assert_array_equal(tf, g["t"][g["kept"]])
    np.testing.assert_array_equal(pf, g["pos"][g["kept"]])
    assert np.random.random() == float(g["rng_after"])


def test_poly_ransac_batched_matches_oracle(gsf):
    """gsf_poly_ransac_dev, many fits in one launch (windows x axes of several tracks), against the numpy restatement of
    sklearn's loop (oracle.poly_ransac_fit) fed the same sample indices; degree 1..3, ragged windows, a window whose
    samples never reach an inlier count of 1 is impossible (the sample itself fits), so status stays 0."""
    from sklearn.utils.random import sample_without_replacement
    from oracle import fusion_oracle as fo
    from oracle.make_golden_gpsfilter import make_track
    rng = np.random.default_rng(5)
    ms, max_trials, thr = 6, 50, 10.0
    for degree in (1, 2, 3):
        ts, ys, widx, off, axes, samples, dyn, dyn_off, want = [], [], [], [0], [], [], [], [], []
        base = 0
        for k in range(6):
            t, pos, _ = make_track(40 + k, n=int(rng.integers(30, 200)), dt=0.1, outlier_frac=0.08, borderline_frac=0.02)
            for axis in range(3):
                nw = len(t)
                rs_dev, rs_ref = np.random.RandomState(1000 * degree + 10 * k + axis), np.random.RandomState(1000 * degree + 10 * k + axis)
                samples.append(np.stack([sample_without_replacement(nw, ms, random_state=rs_dev) for _ in range(max_trials)]))
                mask, used = fo.poly_ransac_fit(t, pos[:, axis], degree, ms, thr, max_trials, rng=rs_ref)
                want.append((mask, used))
                widx.append(base + np.arange(nw)); off.append(off[-1] + nw); axes.append(axis)
                dyn_off.append(sum(len(d) for d in dyn)); dyn.append(gsf.dynamic_max_trials_table(nw, ms, max_trials))
            ts.append(t); ys.append(pos); base += len(t)
        mask, n_trials, status = gsf.poly_ransac(dev(np.concatenate(ts)), dev(np.concatenate(ys)), dev(np.concatenate(widx), torch.int32),
                                                 dev(np.array(off), torch.int64), dev(np.array(axes), torch.int32),
                                                 dev(np.stack(samples), torch.int32), dev(np.concatenate(dyn), torch.int32),
                                                 dev(np.array(dyn_off), torch.int64), ms, degree, max_trials, thr)
        mask, n_trials, status = mask.cpu().numpy().astype(bool), n_trials.cpu().numpy(), status.cpu().numpy()
        assert (status == 0).all()
        for f, (m_ref, used) in enumerate(want):
            np.testing.assert_array_equal(mask[off[f]:off[f + 1]], m_ref, err_msg=f"degree {degree} fit {f}")
            assert n_trials[f] == used


def test_text_parser_and_writer_match_numpy(gsf, tmp_path):
    """N4: gsf_parse_table_dev against np.loadtxt (EKFGPSSLAM.py:110-125, :252-258) -- bit-identical doubles on whitespace-,
    space- and comma-delimited tables with comments, blank lines, exponents, 16-19 digit fields, CRLF, a missing final
    newline -- and gsf_write_pose_rows_dev against np.savetxt with the reference's formats (:1087-1102), byte for byte."""
    import io
    rng = np.random.default_rng(12)
    n = 20011
    tab = np.column_stack([1317384506.0 + np.arange(n) * 0.1037, rng.normal(0, 300, (n, 3)), rng.normal(0, 1, (n, 4))])
    tab[5, 2] = 0.0; tab[6, 3] = -0.0; tab[7, 1] = 1e-12; tab[8, 1] = 123456789012.125
    cases = []
    buf = io.StringIO(); np.savetxt(buf, tab, fmt="%.6f"); cases.append((buf.getvalue(), None))
    buf = io.StringIO(); np.savetxt(buf, tab); cases.append(("# ts x y z qx qy qz qw\n\n" + buf.getvalue().replace("\n", "\r\n", 50), None))      # %.18e
    buf = io.StringIO(); np.savetxt(buf, tab[:300], fmt="%.9f", delimiter=","); cases.append((buf.getvalue().rstrip("\n"), ","))
    buf = io.StringIO(); np.savetxt(buf, tab[:300], fmt="%.12g", delimiter=" "); cases.append((buf.getvalue() + "# tail comment\n", " "))
    cases.append(("1 2 3 # c\n4 5 6\n", None)); cases.append(("7.5\n", None)); cases.append(("1e400 -1e-400 nan -inf 0x\n".replace(" 0x", ""), None))
    for text, delim in cases:
        want = np.loadtxt(io.StringIO(text), delimiter=delim, ndmin=2)
        raw = torch.from_numpy(np.frombuffer(text.encode(), dtype=np.uint8).copy()).cuda()
        got, status = gsf.parse_table(raw, 0 if delim is None else ord(delim), max_cols=want.shape[1])
        assert status == 0, status
        got = got.cpu().numpy()
        assert got.shape == want.shape
        assert np.array_equal(got.view(np.uint64), want.view(np.uint64)) or np.array_equal(got, want, equal_nan=True)
    # errors numpy raises as ValueError
    for text, delim, bit in (("1 2\n3 x\n", None, 1), ("1 2\n3\n", None, 16), ("1  2\n", " ", 1), ("1,2,\n", ",", 1)):
        raw = torch.from_numpy(np.frombuffer(text.encode(), dtype=np.uint8).copy()).cuda()
        _, status = gsf.parse_table(raw, 0 if delim is None else ord(delim), max_cols=4)
        assert status & bit, (text, status)
    # writers
    ts = tab[:, 0]; xyz = tab[:, 1:4] + np.array([455779.0, 5431368.0, 112.0]); quat = tab[:, 4:8]
    xyz[11] = [0.0000005, -0.0000005, 2.5e-7]; xyz[12] = [1.0000005, 2.00000050000001, 0.9999995]; quat[13] = [0.5, -0.125, 1e-9, -1e-9]
    for dec, hdr in (([6, 6, 6, 6, 8, 8, 8, 8], "timestamp x y z qx qy qz qw (UTM)"), ([6, 8, 8, 3, 8, 8, 8, 8], "timestamp lon lat alt qx qy qz qw (WGS84)")):
        ref = io.StringIO()
        np.savetxt(ref, np.column_stack((ts, xyz, quat)), fmt=["%%.%df" % d for d in dec], header=hdr, comments="")
        got = bytes(gsf.write_pose_rows(dev(ts), dev(xyz), dev(quat), dec, hdr + "\n").cpu().numpy())
        assert got == ref.getvalue().encode()


def test_fp32_mode_within_1e_4_m(gsf):
    """Optional fp32 mode (gsf_fuse_batched_f32_dev: fp32 storage relative to fp64 origins and interleaved by 32
    trajectories, fp64 reductions and SVD, fp32 filter in innovation form) against the fp64 oracle on the ORIGINAL data: positions within 1e-4 m (north star), rotation
    and quaternions at fp32 rounding; ragged batch (a partly filled group of 32, lengths 8 .. 1000); a trajectory with an outage is
    handed back with GSF_ST_NEEDS_FP64; the fp64 fused kernel on the same batch agrees to the same tolerance."""
    from gps_optimize_slam_b200 import _lib, synth
    from oracle import fusion_oracle as fo
    trajs = [synth.make_trajectory(300 + k, n=n) for k, n in enumerate((271, 272, 1000, 63, 517, 1000, 8, 271))]
    trajs[4] = synth.make_trajectory(304, n=517, outages=[(100, 140)])
    ts, pos, quat, z, off, offs, maxlen = pack(trajs)
    prm = gsf.params_tensor()
    batch = gsf.to_local_f32(ts, pos, quat, z, off)
    p32, q32, sim3, st = gsf.fuse_batched_f32(batch, prm)
    p, q32 = gsf.from_local_f32(batch, p32, q32)
    p, q32, sim3, st = p.cpu().numpy(), q32.cpu().numpy(), sim3.cpu().numpy(), st.cpu().numpy()
    p64, q64, sim64, st64 = gsf.fuse_batched(ts, pos, quat, z, off, maxlen, prm)
    p64 = p64.cpu().numpy()
    cfg = fo.default_config()
    for b, tr in enumerate(trajs):
        sl = slice(offs[b], offs[b + 1])
        if b == 4:
            assert st[b] == _lib.ST_NEEDS_FP64 and np.isnan(p[sl]).all()
            continue
        assert st[b] == 0, (b, st[b])
        o = oracle_pipeline(tr, cfg)
        assert np.abs(p[sl] - o["pos"]).max() < 1e-4, (b, np.abs(p[sl] - o["pos"]).max())
        assert np.abs(p[sl] - p64[sl]).max() < 1e-4
        assert np.abs(q32[sl] - o["quat"]).max() < 2e-6
        np.testing.assert_allclose(sim3[b, :9].reshape(3, 3), o["R"], rtol=0, atol=2e-6)
        assert abs(sim3[b, 12] - o["s"]) < 2e-6 and np.abs(sim3[b, 9:12] - o["t"]).max() < 2e-2      # t amplifies the rotation's fp32-input rounding by |origin|


def test_dropin_entry_point(gsf, tmp_path):
    """The drop-in module reproduces the reference's run on the shipped pair A (from the
    golden fixture; the reference's files do not travel to the GPU box)."""
    import EKFGPSSLAM as E
    g = load_golden("pairA")
    slam_file, gps_file = tmp_path / "slam.txt", tmp_path / "gnss.txt"
    np.savetxt(slam_file, np.column_stack((g["slam_ts"], g["slam_pos"], g["slam_quat"])))
    np.savetxt(gps_file, g["gnss_raw"], fmt="%.12f")
    E.CONFIG["gps_filtering_ransac"]["enabled"] = True           # the shipped CONFIG: the device pre-filter runs (and removes nothing here)
    np.random.seed(0)
    out = E.main_process(str(slam_file), str(gps_file), save_path=str(tmp_path / "slam_corrected_utm.txt"))
    assert out["utm_zone"] == "39N"
    assert abs(out["s"] - float(g["s"])) < 1e-9
    np.testing.assert_allclose(out["ekf_pos"], g["ekf_pos"], rtol=0, atol=1e-6)
    m, med, rmse, cnt = out["stats"][("primary GPS", "EKF fused")]
    np.testing.assert_allclose([m, med, rmse], g["stats"][2], rtol=1e-7)
    assert cnt == len(g["eval_indices"])
    saved = np.loadtxt(tmp_path / "slam_corrected_utm.txt", skiprows=1)
    assert saved.shape == (271, 8)
    wgs = np.loadtxt(tmp_path / "slam_corrected_wgs84.txt", skiprows=1)
    assert wgs.shape == (271, 8) and abs(wgs[0, 1] - 49.0336) < 1e-3       # "lon" column holds the file's lat (swap)
    R, t, s = E.compute_sim3_transform(g["slam_pos"][:2], g["aligned"][:2])
    assert R is None and t is None and s is None
    dp, dq = E.calculate_relative_pose(g["slam_pos"][3], g["slam_quat"][3], g["slam_pos"][4], g["slam_quat"][4])
    from oracle import fusion_oracle as fo
    dpo, dqo = fo.relative_pose(g["slam_pos"][3], g["slam_quat"][3], g["slam_pos"][4], g["slam_quat"][4])
    np.testing.assert_allclose(dp, dpo, atol=1e-12); np.testing.assert_allclose(dq, dqo, atol=1e-12)


def test_relative_pose_many_pairs(gsf):
    """calculate_relative_pose (EKFGPSSLAM.py:77-92) through the drop-in on many random pose pairs -- unnormalised
    quaternions (scipy normalises on construction), large UTM-scale positions, a half-turn -- and the zero-quaternion
    fallback (:84-86: zero motion, identity rotation) against the oracle (scipy Rotation)."""
    import EKFGPSSLAM as E
    from oracle import fusion_oracle as fo
    rng = np.random.default_rng(31)
    for k in range(60):
        q1 = rng.normal(size=4) * rng.choice([1.0, 0.3, 7.0]); q2 = rng.normal(size=4) * rng.choice([1.0, 2.0])
        if k == 7:
            q2 = -q1.copy()                                  # same rotation, opposite sign
        if k == 8:
            q1 = np.array([0.0, 0.0, 1.0, 0.0]); q2 = np.array([0.0, 0.0, 0.0, 1.0])       # half-turn about z
        p1 = rng.normal(size=3) * rng.choice([1.0, 5e6]); p2 = p1 + rng.normal(size=3) * rng.choice([0.01, 1.0, 50.0])
        dp, dq = E.calculate_relative_pose(p1, q1, p2, q2)
        dpo, dqo = fo.relative_pose(p1, q1, p2, q2)
        np.testing.assert_allclose(dp, dpo, rtol=0, atol=1e-9 * max(1.0, np.abs(p2 - p1).max()))
        assert min(np.abs(dq - dqo).max(), np.abs(dq + dqo).max()) < 1e-12 and np.abs(dq - dqo).max() < 1e-12
    for q1, q2 in ((np.zeros(4), np.array([0.0, 0.0, 0.0, 1.0])), (np.array([0.1, 0.2, 0.3, 0.9]), np.zeros(4))):
        dp, dq = E.calculate_relative_pose(np.zeros(3), q1, np.ones(3), q2)
        dpo, dqo = fo.relative_pose(np.zeros(3), q1, np.ones(3), q2)
        np.testing.assert_array_equal(dp, dpo); np.testing.assert_array_equal(dq, dqo)


# ----------------------------------------------------------------------------- long trajectories (no length limit)
def _nn_errors_chunked(traj, cand):
    """cdist(...).min(axis=1) of EKFGPSSLAM.py:1030-1031 in row blocks (the full matrix of a 20 000-pose
    trajectory would need gigabytes)."""
    from scipy.spatial import distance
    return np.concatenate([distance.cdist(traj[a:a + 1024], cand, "euclidean").min(axis=1) for a in range(0, len(traj), 1024)])


@pytest.mark.parametrize("n,outages", [(4541, ()), (4541, [(2100, 2400), (4000, 4075)]), (20000, ()),
                                       (20000, [(0, 80), (2250, 2400), (9000, 9700), (19900, 20000)])])
def test_long_trajectory_through_dropin_matches_oracle(gsf, n, outages):
    """Trajectories beyond the shared-memory staging (KITTI-00 length and 20 000 poses, with and without GNSS
    outages that span tile boundaries, start at pose 0 or run to the end) through the drop-in's
    apply_ekf_correction / evaluate_errors and through gsf_fuse_batched_dev, against the oracle."""
    import contextlib, io
    import EKFGPSSLAM as E
    from gps_optimize_slam_b200 import synth
    from oracle import fusion_oracle as fo
    cfg = fo.default_config()
    tr = synth.make_trajectory(900 + n // 1000 + len(outages), n=n, dt=0.1, speed=10.0, outages=outages)
    valid_rows = ~np.isnan(tr["gps"]).any(1)
    slam = {"timestamps": tr["ts"], "positions": tr["pos"], "quaternions": tr["quat"]}
    gps = {"timestamps": tr["ts"][valid_rows], "positions": tr["gps"][valid_rows]}
    # oracle: the reference's own order of steps (association -> selection -> Umeyama -> apply -> EKF -> evaluation)
    aligned, valid = fo.associate(tr["ts"], gps["timestamps"], gps["positions"])
    assert np.array_equal(valid, valid_rows)
    sel = fo.sim3_point_selection(tr["ts"], valid)
    R, t, s = fo.umeyama(tr["pos"][sel], aligned[sel])
    sp, sq = fo.sim3_apply(tr["pos"], tr["quat"], R, t, s)
    fp, fq = fo.ekf_fuse(tr["ts"], tr["pos"], tr["quat"], aligned, valid, sp[0], sq[0], cfg)
    ev = fo.evaluation_indices(tr["ts"], valid)
    with contextlib.redirect_stdout(io.StringIO()):
        got_p, got_q = E.apply_ekf_correction(slam, gps, sp, sq, E.CONFIG)
        m, med, rmse, cnt = E.evaluate_errors(got_p, aligned, tr["ts"])
    # the evaluation is checked on the same trajectory bits (one ulp of a UTM coordinate is 9.3e-10 m, i.e. 5e-9 of a 0.2 m error)
    want = fo.error_stats(_nn_errors_chunked(got_p[ev], aligned[ev]))
    np.testing.assert_allclose(got_p, fp, rtol=0, atol=POS_ATOL)
    np.testing.assert_allclose(got_q, fq, rtol=0, atol=ROT_ATOL)
    assert cnt == len(ev)
    np.testing.assert_allclose([m, med, rmse], want, rtol=1e-10)
    # the fused entry point (selection + Umeyama inside), alone and in a batch with short trajectories
    z = np.where(valid[:, None], aligned, np.nan)
    long_tr = dict(ts=tr["ts"], pos=tr["pos"], quat=tr["quat"], gps=z)
    short = [synth.make_trajectory(5, n=1000, dt=0.1, speed=10.0, outages=[(300, 420)]), synth.make_trajectory(1, n=271)]
    batch = [short[0], long_tr, short[1]]
    ts_d, pos_d, quat_d, z_d, off_d, off, max_len = pack(batch)
    p, q, sim3, st = gsf.fuse_batched(ts_d, pos_d, quat_d, z_d, off_d, max_len, gsf.params_tensor(cfg))
    p, q, sim3, st = p.cpu().numpy(), q.cpu().numpy(), sim3.cpu().numpy(), st.cpu().numpy()
    assert (st == 0).all(), st
    sl = slice(off[1], off[2])
    assert int(sim3[1, 13]) == len(sel) and int(sim3[1, 14]) == int(valid.sum())
    np.testing.assert_allclose(sim3[1, :9].reshape(3, 3), R, atol=ROT_ATOL)
    np.testing.assert_allclose(sim3[1, 9:12], t, rtol=1e-12, atol=POS_ATOL)
    assert abs(sim3[1, 12] - s) < 1e-11
    np.testing.assert_allclose(p[sl], fp, rtol=0, atol=POS_ATOL)
    np.testing.assert_allclose(q[sl], fq, rtol=0, atol=ROT_ATOL)
    for b in (0, 2):
        o = oracle_pipeline(batch[b], cfg)
        np.testing.assert_allclose(p[off[b]:off[b + 1]], o["pos"], rtol=0, atol=POS_ATOL)
    # the long-trajectory kernel is deterministic and agrees with the literal one-thread recursion
    p2, q2, _, _ = gsf.fuse_batched(ts_d, pos_d, quat_d, z_d, off_d, max_len, gsf.params_tensor(cfg))
    assert np.array_equal(p2.cpu().numpy(), p) and np.array_equal(q2.cpu().numpy(), q)
    one = dev(np.array([0, n]), torch.int64)
    ps, qs, _ = gsf.ekf_strict_batched(dev(tr["ts"]), dev(tr["pos"]), dev(tr["quat"]), dev(z), one, gsf.params_tensor(cfg), dev(sp[:1]), dev(sq[:1]))
    np.testing.assert_allclose(ps.cpu().numpy(), fp, rtol=0, atol=POS_ATOL)


def test_long_trajectory_sharp_turn_and_window_fallbacks(gsf):
    """Long trajectory whose first gap-free run is too short (selection falls back to all valid points), with a sharp
    turn inside an outage (RTS skipped, blended update) -- the rare branches of :972-998 and :879-928 in the tiled kernel."""
    from gps_optimize_slam_b200 import synth
    from oracle import fusion_oracle as fo
    for steps in (0, 4):
        cfg = fo.default_config()
        cfg["rts_decision"]["default_ekf_transition_steps_on_sharp_turn"] = steps
        tr = synth.make_trajectory(950, n=6000, dt=0.1, speed=10.0, outages=[(3, 80), (2290, 2420), (5000, 5100)], sharp_turn_at=2350)
        o = oracle_pipeline(tr, cfg)
        ts_d, pos_d, quat_d, z_d, off_d, off, max_len = pack([tr])
        p, q, sim3, st = gsf.fuse_batched(ts_d, pos_d, quat_d, z_d, off_d, max_len, gsf.params_tensor(cfg))
        assert int(st.cpu()[0]) == 0
        assert int(sim3.cpu()[0, 13]) == len(o["sel"])
        np.testing.assert_allclose(sim3.cpu().numpy()[0, :9].reshape(3, 3), o["R"], atol=ROT_ATOL)
        np.testing.assert_allclose(p.cpu().numpy(), o["pos"], rtol=0, atol=POS_ATOL)
        np.testing.assert_allclose(q.cpu().numpy(), o["quat"], rtol=0, atol=ROT_ATOL)


def test_ate_kernel_any_length_and_far_queries(gsf):
    """The bucketed nearest-neighbour search against brute force: a 20 000-pose loop (candidates from other laps are
    nearer than the index-matched one), queries far from every candidate (raw SLAM frame vs UTM: no pruning possible),
    duplicate errors around the median, a track along y (bins along the other axis), NaN rows and a 1-point set."""
    from oracle import fusion_oracle as fo
    rng = np.random.default_rng(12)
    n = 20000
    lap = 2 * np.pi * np.arange(n) / 5000.0
    cand = np.stack([455000 + 300 * np.cos(lap), 5431000 + 900 * np.sin(lap), 100 + 0.01 * np.arange(n) % 7], axis=1) + rng.normal(0, 0.3, (n, 3))
    cand[rng.uniform(size=n) < 0.1] = np.nan
    ts = np.arange(n) * 0.1
    traj = np.where(np.isnan(cand), 0.0, cand) + rng.normal(0, 0.5, (n, 3))
    far = rng.normal(0, 100.0, (n, 3))
    dup = np.where(np.isnan(cand), 0.0, cand) + np.array([0.25, 0.0, 0.0])
    one = dev(np.array([0, n]), torch.int64)
    valid = ~np.isnan(cand).any(1)
    ev = fo.evaluation_indices(ts, valid)
    for t_ in (traj, far, dup):
        got = gsf.ate_nn_batched(dev(t_), dev(cand), dev(ts), one, n, 5.0).cpu().numpy()[0]
        want = fo.error_stats(_nn_errors_chunked(t_[ev], cand[ev]))
        assert int(got[3]) == len(ev)
        np.testing.assert_allclose(got[:3], want, rtol=1e-10)
    # ragged batch: shared-memory variant, sizes 1 .. 6000, one trajectory without any evaluation point
    lens = [1, 60, 61, 1000, 6000, 333]
    off = np.concatenate(([0], np.cumsum(lens))).astype(np.int64)
    c = cand[: off[-1]].copy(); tr2 = traj[: off[-1]].copy()
    tsb = np.concatenate([np.arange(k) * 0.1 for k in lens])
    c[off[3]:off[4], :2] = c[off[3]:off[4], 1::-1] * [1e-3, 1.0]          # a track along y with a tiny x extent
    got = gsf.ate_nn_batched(dev(tr2), dev(c), dev(tsb), dev(off, torch.int64), max(lens), 5.0).cpu().numpy()
    for b, k in enumerate(lens):
        sl = slice(off[b], off[b + 1])
        v = ~np.isnan(c[sl]).any(1)
        evb = fo.evaluation_indices(tsb[sl], v)
        assert int(got[b, 3]) == len(evb)
        if len(evb):
            np.testing.assert_allclose(got[b, :3], fo.error_stats(_nn_errors_chunked(tr2[sl][evb], c[sl][evb])), rtol=1e-10)
        else:
            assert np.isnan(got[b, :3]).all()


def test_dropin_main_process_pair_b(gsf, tmp_path):
    """Shipped pair B (yolotum04.txt + combined_output.txt: zone 32N, 270/271 valid, RTS over [0, 1]) through the
    drop-in's main_process against the unmodified reference's outputs (tests/golden/pairB.npz)."""
    import contextlib, io
    import EKFGPSSLAM as E
    g = load_golden("pairB")
    slam_file, gps_file = tmp_path / "slam.txt", tmp_path / "gnss.txt"
    np.savetxt(slam_file, np.column_stack((g["slam_ts"], g["slam_pos"], g["slam_quat"])), fmt="%.18e")
    np.savetxt(gps_file, g["gnss_raw"], fmt="%.18e")
    E.CONFIG["gps_filtering_ransac"]["enabled"] = True           # the shipped CONFIG: the device pre-filter runs (and removes nothing here)
    np.random.seed(0)
    with contextlib.redirect_stdout(io.StringIO()):
        out = E.main_process(str(slam_file), str(gps_file))
    assert out["utm_zone"] == str(g["utm_zone"]) == "32N"
    assert int(out["valid"].sum()) == int(g["valid"].sum()) == 270
    np.testing.assert_allclose(out["R"], g["R"], rtol=0, atol=ROT_ATOL)
    assert abs(out["s"] - float(g["s"])) < 1e-11
    np.testing.assert_allclose(out["sim3_pos"], g["sim3_pos"], rtol=0, atol=POS_ATOL)
    np.testing.assert_allclose(out["ekf_pos"], g["ekf_pos"], rtol=0, atol=POS_ATOL)
    np.testing.assert_allclose(out["ekf_quat"], g["ekf_quat"], rtol=0, atol=ROT_ATOL)
    for k, label in enumerate(("raw SLAM", "Sim3 aligned", "EKF fused")):
        m, med, rmse, cnt = out["stats"][("primary GPS", label)]
        # errors are differences of UTM-scale coordinates computed from trajectories that agree to POS_ATOL: same absolute budget
        np.testing.assert_allclose([m, med, rmse], g["stats"][k], rtol=0, atol=POS_ATOL)
        assert cnt == len(g["eval_indices"])
