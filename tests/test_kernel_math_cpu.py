"""CPU checks of what the CUDA kernels compute:
  * the kernels' scalar math (gsf_common.cuh / gsf_ekf_strict.cuh) compiled with g++ and
    compared with scipy / the oracle (tests/hostmath);
  * the kernel's time-parallel formulation (oracle/kernel_model.py: telescoped odometry,
    Moebius + affine scans, closed-form RTS) compared with the step-by-step oracle."""
import ctypes
import warnings

import numpy as np
import pytest
from scipy.spatial.transform import Rotation

from conftest import GOLDEN_CASES, load_golden
from gps_optimize_slam_b200 import config, synth
from oracle import fusion_oracle as fo, kernel_model as km

DP = ctypes.POINTER(ctypes.c_double)


def P(a):
    return a.ctypes.data_as(DP)


def test_yaw_and_from_matrix_match_scipy(hostmath):
    rng = np.random.default_rng(0)
    quats = rng.normal(size=(500, 4))
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        gimbal = [Rotation.from_euler("zyx", [a, s * np.pi / 2, b]).as_quat()
                  for a, b, s in zip(rng.uniform(-3, 3, 40), rng.uniform(-3, 3, 40), [1, -1] * 20)]
        for q in list(quats) + gimbal:
            q = np.ascontiguousarray(q, dtype=float)
            assert abs(hostmath.hm_yaw_zyx(P(q)) - Rotation.from_quat(q).as_euler("zyx")[0]) < 1e-13
    for i in range(300):
        R = np.ascontiguousarray(Rotation.random(random_state=i).as_matrix())
        q = np.zeros(4)
        hostmath.hm_quat_from_matrix(P(R), P(q))
        np.testing.assert_allclose(q, Rotation.from_matrix(R).as_quat(), atol=1e-15)


def _umeyama_host(hostmath, src, dst):
    n = len(src)
    ms, md = src.mean(0), dst.mean(0)
    a, b = src - ms, dst - md
    H = np.ascontiguousarray(a.T @ b)
    R, t, s = np.zeros(9), np.zeros(3), ctypes.c_double()
    st = hostmath.hm_umeyama_finish(n, P(ms), P(md), P(H), ctypes.c_double(float((a * a).sum())), P(R), P(t), ctypes.byref(s))
    return R.reshape(3, 3), t, s.value, st


@pytest.mark.parametrize("seed", range(8))
def test_jacobi_umeyama_matches_lapack_path(hostmath, seed):
    tr = synth.make_trajectory(2000 + seed)
    src, dst = tr["pos"], tr["gps"]
    if seed >= 4:
        dst = dst[:, [1, 0, 2]]                       # mirrored target: reflection branch
    R, t, s = fo.umeyama(src, dst)
    R2, t2, s2, st = _umeyama_host(hostmath, src, dst)
    assert st == 0
    np.testing.assert_allclose(R2, R, atol=1e-9)
    np.testing.assert_allclose(t2, t, rtol=1e-12, atol=1e-7)
    assert abs(s2 - s) < 1e-12
    assert abs(np.linalg.det(R2) - 1) < 1e-12


def test_umeyama_golden_pairs(hostmath):
    for case in ("pairA", "pairB"):
        g = load_golden(case)
        sel = g["sim3_indices"]
        R2, t2, s2, st = _umeyama_host(hostmath, g["slam_pos"][sel], g["aligned"][sel])
        np.testing.assert_allclose(R2, g["R"], atol=1e-9)
        np.testing.assert_allclose(t2, g["t"], rtol=1e-12)
        assert abs(s2 - float(g["s"])) < 1e-12 and st == 0


def test_umeyama_collinear_flags_degenerate(hostmath):
    src = np.outer(np.arange(10.0), [1.0, 2.0, 3.0])
    _, _, _, st = _umeyama_host(hostmath, src, 2 * src + 1)
    assert st == 2


def _strict(hostmath, ts, p, q, z, ip, iq, cfg):
    prm = config.pack_fuse_params(cfg)
    op, oq = np.zeros_like(p), np.zeros_like(q)
    fn = hostmath.hm_ekf_strict
    fn.argtypes = [ctypes.c_long] + [DP] * 6 + [ctypes.c_void_p] + [DP] * 2
    st = fn(len(ts), P(ts), P(p), P(q), P(z), P(np.ascontiguousarray(ip)), P(np.ascontiguousarray(iq)),
            prm.ctypes.data, P(op), P(oq))
    return op, oq, st


@pytest.mark.parametrize("case", GOLDEN_CASES)
def test_strict_recursion_matches_golden(hostmath, case):
    g = load_golden(case)
    arrs = [np.ascontiguousarray(g[k]) for k in ("slam_ts", "slam_pos", "slam_quat", "aligned")]
    op, oq, st = _strict(hostmath, *arrs, g["sim3_pos"][0], g["sim3_quat"][0], fo.default_config())
    assert st == 0
    np.testing.assert_allclose(op, g["ekf_pos"], rtol=0, atol=2e-8)
    np.testing.assert_allclose(oq, g["ekf_quat"], rtol=0, atol=1e-13)


def test_strict_recursion_zero_quaternion_and_blend(hostmath):
    cfg = fo.default_config()
    cfg["rts_decision"]["default_ekf_transition_steps_on_sharp_turn"] = 4
    tr = synth.make_trajectory(1003, outages=[(50, 90)], sharp_turn_at=70)
    ts, p, q, z = [np.ascontiguousarray(tr[k]) for k in ("ts", "pos", "quat", "gps")]
    valid = ~np.isnan(z).any(1)
    R, t, s = fo.umeyama(p[valid], z[valid])
    sp, sq = fo.sim3_apply(p[:1], q[:1], R, t, s)
    q[120] = 0; q[60] = 0; q[61] = 0
    fp, fq = fo.ekf_fuse(ts, p, q, z, valid, sp[0], sq[0], cfg)
    op, oq, st = _strict(hostmath, ts, p, q, z, sp[0], sq[0], cfg)
    assert st == 4
    np.testing.assert_allclose(op, fp, atol=2e-8, rtol=0)
    np.testing.assert_allclose(oq, fq, atol=1e-13, rtol=0)


CASES = {
    "clean": dict(seed=1000),
    "outage": dict(seed=1001, outages=[(50, 90)]),
    "outage_at_start": dict(seed=1002, outages=[(0, 7), (100, 101), (200, 260)]),
    "sharp": dict(seed=1003, outages=[(50, 90)], sharp_turn_at=70),
    "ends_in_outage": dict(seed=1004, outages=[(250, 271)]),
    "long": dict(seed=1005, n=1000, dt=0.1, speed=10, outages=[(300, 420)]),
    "selection_gap": dict(seed=1006, n=400, outages=[(100, 160)]),
    "unnormalised_quats": dict(seed=1007, quat_scale_jitter=1e-3),
}


@pytest.mark.parametrize("name", CASES)
@pytest.mark.parametrize("threads", [32, 128])
def test_kernel_formulation_matches_oracle(name, threads):
    tr = synth.make_trajectory(**CASES[name])
    ts, p, q, z = tr["ts"], tr["pos"], tr["quat"], tr["gps"]
    cfg = fo.default_config()
    if name == "sharp":
        cfg["rts_decision"]["default_ekf_transition_steps_on_sharp_turn"] = 4
    valid = ~np.isnan(z).any(1)
    sel = fo.sim3_point_selection(ts, valid)
    R, t, s = fo.umeyama(p[sel], z[sel])
    sp, sq = fo.sim3_apply(p, q, R, t, s)
    fp, fq, events = fo.ekf_fuse(ts, p, q, z, valid, sp[0], sq[0], cfg, return_events=True)
    m = km.fused_model(ts, p, q, z, cfg, T=threads)
    assert m["sel_count"] == len(sel)
    assert [(a, b - 1, c) for a, b, c in events] == m["segments"]
    np.testing.assert_allclose(m["R"], R, atol=1e-12)
    np.testing.assert_allclose(m["pos"], fp, rtol=0, atol=5e-8)
    np.testing.assert_allclose(m["quat"], fq, rtol=0, atol=1e-13)


def test_text_conversions_match_python(hostmath):
    """gsf_text.cuh compiled for the host: parse_double against Python's float() (strtod-quality: Clinger fast path +
    Eisel-Lemire) and format_fixed against printf's "%.Df" -- the conversions behind gsf_parse_table_dev /
    gsf_write_pose_rows_dev (np.loadtxt / np.savetxt of EKFGPSSLAM.py:110-125, :252-258, :1087-1102)."""
    import ctypes
    import random
    hostmath.hm_parse_double.argtypes = [ctypes.c_char_p, ctypes.c_int, ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_int)]
    hostmath.hm_format_fixed.argtypes = [ctypes.c_double, ctypes.c_int, ctypes.c_char_p, ctypes.POINTER(ctypes.c_int)]
    rnd = random.Random(3)
    out, flag = ctypes.c_double(), ctypes.c_int(0)
    for i in range(30000):
        k = rnd.random()
        if k < 0.3:
            s = "%.6f" % rnd.uniform(-2e9, 2e9)
        elif k < 0.55:
            s = repr(rnd.uniform(-1, 1) * 10 ** rnd.randint(-30, 30))
        elif k < 0.8:
            s = "%.18e" % (rnd.uniform(-1, 1) * 10 ** rnd.randint(-300, 300))
        else:
            s = str(rnd.getrandbits(63)) + "e" + str(rnd.randint(-330, 290))
        flag.value = 0
        used = hostmath.hm_parse_double(s.encode(), len(s), ctypes.byref(out), ctypes.byref(flag))
        assert used == len(s) and flag.value == 0 and out.value == float(s), s
    for s, want in (("nan", None), ("-inf", -np.inf), ("Infinity", np.inf), ("1e400", np.inf), ("5e-324", 5e-324), ("-0", -0.0), ("+.5e1", 5.0)):
        used = hostmath.hm_parse_double(s.encode(), len(s), ctypes.byref(out), ctypes.byref(flag))
        assert used == len(s) and (np.isnan(out.value) if want is None else out.value == want)
    assert hostmath.hm_parse_double(b"abc", 3, ctypes.byref(out), ctypes.byref(flag)) == 0
    buf = ctypes.create_string_buffer(64)
    for i in range(30000):
        k = rnd.random()
        if k < 0.4:
            x = rnd.uniform(-2e9, 2e9)
        elif k < 0.6:
            x = rnd.uniform(-1, 1) * 10 ** rnd.randint(-12, 3)
        elif k < 0.8:
            x = round(rnd.uniform(-1e7, 1e7), rnd.choice([3, 6, 8])) + rnd.choice([0, 5e-4, 5e-7, 5e-9, -5e-9])
        else:
            x = rnd.randint(-10 ** 9, 10 ** 9) / rnd.choice([2, 4, 8, 16, 1024, 2 ** 20, 2 ** 30])
        D = rnd.choice([3, 6, 8])
        flag.value = 0
        n = hostmath.hm_format_fixed(x, D, buf, ctypes.byref(flag))
        assert buf.raw[:n].decode() == ("%." + str(D) + "f") % x and flag.value == 0, x
    for x in (0.0, -0.0, float("nan"), float("inf"), -float("inf"), 0.5, 1.5, 2.5, 9.9999999999, 999999.9999995):
        n = hostmath.hm_format_fixed(x, 6, buf, ctypes.byref(flag))
        assert buf.raw[:n].decode() == "%.6f" % x


def test_local_halo_spline_solve_matches_global_solve():
    """The algorithm of csrc/gsf_assoc_long.cu in numpy (oracle/assoc_long_model.py: 13-knot chunks, 20-knot halos, forward
    elimination + elimination from the far end joined at the chunk's last row) against the global not-a-knot solve per
    segment (scipy CubicSpline = the spline behind interp1d(kind='cubic'), EKFGPSSLAM.py:351-380): irregular spacing, gaps
    that cut segments of 3, 1, 4 and hundreds of knots.  The halo bound is 0.268^20 = 3.7e-12 of the moments' size; a
    6-knot halo shows the decay it rests on."""
    from scipy.interpolate import CubicSpline
    from oracle.assoc_long_model import local_halo_moments
    rng = np.random.default_rng(3)
    M = 700
    steps = np.minimum(rng.uniform(0.02, 0.4, M) * rng.choice([1.0, 1.0, 5.0], M), 4.0)
    for k in (100, 103, 104, 108, 400, 640):
        steps[k] = rng.uniform(5.5, 8.0)
    t = 10 + np.cumsum(steps)
    y = 5e6 + 8 * t + 30 * np.sin(t / 7) + rng.normal(0, 0.3, M)
    h = np.diff(t)
    ss = np.concatenate([[0], np.nonzero(h > 5)[0] + 1]); se = np.concatenate([ss[1:] - 1, [M - 1]])

    def worst(H):
        m = local_halo_moments(t, y, gap=5.0, CH=13, H=H)
        w = 0.0
        for a, b in zip(ss, se):
            if b - a + 1 < 4:
                assert np.isnan(m[a:b + 1]).all()                      # linear / skipped segments: no moments
                continue
            mg = CubicSpline(t[a:b + 1], y[a:b + 1], bc_type="not-a-knot")(t[a:b + 1], 2)
            assert np.isnan(m[a]) and np.isnan(m[b])                   # end moments are formed at evaluation time
            w = max(w, np.abs(m[a + 1:b] - mg[1:-1]).max() / np.abs(mg).max())
        return w

    assert worst(20) < 1e-11
    assert 1e-9 < worst(6) < 0.268 ** 6 * 4
