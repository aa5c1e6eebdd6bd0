"""No-GPU checks: the C-ABI library loads and exports every symbol include/gsf.h declares, the
product refuses to run without a device, and the (unpinned) UTM oracle passes its
known-answer checks."""
import os
import re

import numpy as np
import pytest

from conftest import ROOT
from oracle import utm_kruger as uk


@pytest.fixture(scope="module")
def lib():
    from gps_optimize_slam_b200 import _lib, build
    build.build()                       # nvcc cross-compiles sm_100a without a GPU
    return _lib.load()


def test_header_symbols_exported(lib):
    from gps_optimize_slam_b200 import _lib
    header = open(os.path.join(ROOT, "include", "gsf.h")).read()
    declared = set(re.findall(r"\b(gsf_[a-z0-9_]+)\s*\(", header))
    assert declared, "no declarations parsed"
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    for name in declared:
        assert hasattr(lib, name), name
    assert b"sm_100a" in lib.gsf_version()


def test_params_record_size():
    from gps_optimize_slam_b200 import config
    blob = config.pack_fuse_params()
    assert blob.nbytes == 184 and blob.dtype == np.uint8
    vals = np.frombuffer(blob.tobytes()[:176], dtype=np.float64)
    assert vals[0] == 0.1 and vals[9] == 0.7 and vals[14] == 0.2 and vals[17] == 5.0 and vals[18] == 180.0
    assert abs(vals[19] - np.deg2rad(45.0)) < 1e-15 and vals[20] == 4.0 and vals[21] == 5.0
    assert np.frombuffer(blob.tobytes()[176:], dtype=np.int32).tolist() == [4, 0]


def test_no_cpu_fallback(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from gps_optimize_slam_b200 import _lib, fusion
    assert lib.gsf_device_sm_count() == _lib.GSF_E_NO_DEVICE
    assert b"no CPU fallback" in lib.gsf_last_error()
    t = torch.zeros(4, dtype=torch.float64)
    with pytest.raises(_lib.GsfError):
        fusion.utm_forward(t, t, 32, False)
    z = np.zeros((4, 3)); off = np.array([0, 4], dtype=np.int64)
    from gps_optimize_slam_b200.config import pack_fuse_params
    with pytest.raises(_lib.GsfError):
        fusion.fuse_batched_host(np.zeros(4), z, np.zeros((4, 4)), z, off, 4, pack_fuse_params())


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "gps_optimize_slam_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, re.M), f
                assert "from .. import oracle" not in text and "from oracle" not in text, f
    entry = open(os.path.join(ROOT, "EKFGPSSLAM.py")).read()
    assert not re.search(r"^\s*(from|import)\s+oracle\b", entry, re.M)


# ---- UTM oracle known answers (closed-form properties; pyproj itself is unavailable)
def test_utm_central_meridian_and_equator():
    e, n = uk.utm_forward(np.array([9.0]), np.array([0.0]), 32, False)
    assert abs(e[0] - 500000.0) < 1e-9 and abs(n[0]) < 1e-9
    # northing on the central meridian = k0 * meridian arc (numerical quadrature of M(phi))
    from scipy.integrate import quad
    a, f = uk.WGS84_A, uk.WGS84_F
    e2 = f * (2 - f)
    for lat in (10.0, 49.0, 80.0):
        arc, _ = quad(lambda p: a * (1 - e2) / (1 - e2 * np.sin(p) ** 2) ** 1.5, 0, np.deg2rad(lat), epsabs=1e-7, epsrel=1e-14)
        _, n = uk.utm_forward(np.array([9.0]), np.array([lat]), 32, False)
        assert abs(n[0] - uk.UTM_K0 * arc) < 2e-6
    # pole-to-equator arc: 10 001 965.729 m (WGS84)
    _, n = uk.utm_forward(np.array([9.0]), np.array([89.999999]), 32, False)
    assert abs(n[0] / uk.UTM_K0 - 10001965.729) < 0.2


def test_utm_published_point_and_south():
    # CN Tower 43.642567 N, 79.387139 W -> zone 17N, E 630084 N 4833439 (published to 1 m)
    e, n = uk.utm_forward(np.array([-79.387139]), np.array([43.642567]), 17, False)
    assert abs(e[0] - 630084) < 1.5 and abs(n[0] - 4833439) < 1.5
    e1, n1 = uk.utm_forward(np.array([151.2]), np.array([-33.85]), 56, True)
    e2, n2 = uk.utm_forward(np.array([151.2]), np.array([33.85]), 56, False)
    assert abs(e1[0] - e2[0]) < 1e-9 and abs((1e7 - n1[0]) - n2[0]) < 1e-8


def test_utm_roundtrip_and_conformality():
    rng = np.random.default_rng(1)
    lon = 9.0 + rng.uniform(-3.5, 3.5, 2000)
    lat = rng.uniform(-80, 84, 2000)
    e, n = uk.utm_forward(lon, lat, 32, False)
    lon2, lat2 = uk.utm_inverse(e, n, 32, False)
    assert np.abs(lon2 - lon).max() < 1e-12 and np.abs(lat2 - lat).max() < 1e-12
    # Cauchy-Riemann in isometric coordinates: the map must be conformal
    h = 1e-6
    lat0, lon0 = np.deg2rad(49.0), np.deg2rad(1.5)
    f = lambda la, lo: np.array(uk.utm_forward(np.rad2deg([lo + np.deg2rad(9)]), np.rad2deg([la]), 32, False)).ravel()
    d_lat = (f(lat0 + h, lon0) - f(lat0 - h, lon0)) / (2 * h)
    d_lon = (f(lat0, lon0 + h) - f(lat0, lon0 - h)) / (2 * h)
    e2 = uk.WGS84_F * (2 - uk.WGS84_F)
    # d(isometric lat)/d(lat) = (1-e2)/((1-e2 sin^2) cos)
    dpsi = (1 - e2) / ((1 - e2 * np.sin(lat0) ** 2) * np.cos(lat0))
    d_psi = d_lat / dpsi
    assert abs(d_psi[1] - d_lon[0]) / abs(d_lon[0]) < 1e-6 and abs(d_psi[0] + d_lon[1]) / abs(d_lon[0]) < 1e-6


def test_zone_rule_and_mask():
    assert uk.utm_zone_from_means(np.array([8.39]), np.array([49.0])) == (32, False)
    assert uk.utm_zone_from_means(np.array([49.03]), np.array([8.39])) == (39, False)
    assert uk.utm_zone_from_means(np.array([-179.9]), np.array([-1.0])) == (1, True)
    m = uk.gnss_validity_mask(np.array([0.0, 91.0, 45.0, 45.0]), np.array([10.0, 10.0, 0.0, 181.0]))
    assert not m.any()
