"""UTM pin that does not depend on the recalled Krueger coefficient table: ``oracle/utm_mp.py`` evaluates the
definition of the ellipsoidal transverse Mercator projection with mpmath at 40 digits (series coefficients by
quadrature of the exact latitude functions, order 12) and wrote tests/golden/utm_mp.npz.  The fp64 restatement
(oracle/utm_kruger.py, what the GPU kernels are compared with) must agree with it to 2 ulp of the output
coordinate (one ulp of a 9.3e6 m northing is 1.9e-9 m).  The reference's own projection is pyproj/PROJ
(EKFGPSSLAM.py:266-271, :291-296), absent from this image."""
import os

import numpy as np
import pytest

from oracle import utm_kruger as uk

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "utm_mp.npz")
TOL_M = 4e-9          # 2 ulp at the largest northing of the lattice
TOL_DEG = 1e-13


def _groups():
    d = np.load(GOLD)
    P = d["points"]
    for zone in np.unique(P[:, 2]):
        for south in (0.0, 1.0):
            m = (P[:, 2] == zone) & (P[:, 3] == south)
            if m.any():
                yield int(zone), bool(south), P[m]


def test_forward_matches_40_digit_evaluation():
    worst = 0.0
    count = 0
    for zone, south, P in _groups():
        e, n = uk.utm_forward(P[:, 0], P[:, 1], zone, south)
        worst = max(worst, np.abs(e - P[:, 4]).max(), np.abs(n - P[:, 5]).max())
        count += len(P)
    assert count >= 800 and worst < TOL_M, worst


def test_inverse_matches_40_digit_evaluation():
    for zone, south, P in _groups():
        lon, lat = uk.utm_inverse(P[:, 4], P[:, 5], zone, south)
        assert np.abs(lon - P[:, 0]).max() < TOL_DEG and np.abs(lat - P[:, 1]).max() < TOL_DEG


def test_coefficient_table_matches_quadrature_derivation():
    """the table of oracle/utm_kruger.py (and of csrc/gsf_capi.cu, tested against it on the GPU) vs the coefficients
    stored with the fixture: differences are the n^7 truncation of the published 6th-order series."""
    d = np.load(GOLD)
    n = uk.WGS84_F / (2 - uk.WGS84_F)
    assert abs(uk._A / uk.WGS84_A - float(d["A_over_a"])) < 2e-16
    for j in range(6):
        assert abs(uk._ALPHA[j] - d["alpha"][j]) < 4 * n ** 7, j
        assert abs(uk._BETA[j] - d["beta"][j]) < 4 * n ** 7, j
    # what the truncation leaves out (orders 7..12) moves a point by less than a picometre
    assert np.abs(d["alpha"][6:]).sum() * uk._A * np.cosh(2 * 7 * 0.07) < 1e-12 * 1e3


def test_fixture_regenerates_from_mpmath():
    """guards the committed fixture: re-derive the series and re-evaluate a few lattice points here."""
    mpmath = pytest.importorskip("mpmath")
    from oracle import utm_mp
    A_over_a, alpha, beta = utm_mp.coefficients(order=8, dps=30)
    d = np.load(GOLD)
    assert abs(float(A_over_a) - float(d["A_over_a"])) < 1e-16
    np.testing.assert_allclose([float(a) for a in alpha[:6]], d["alpha"][:6], rtol=1e-9, atol=1e-24)
    np.testing.assert_allclose([float(b) for b in beta[:6]], d["beta"][:6], rtol=1e-9, atol=1e-24)
    P = d["points"]
    for row in P[:: max(1, len(P) // 7)]:
        E, N = utm_mp.utm_forward_mp(row[0], row[1], int(row[2]), bool(row[3]), order=8, dps=30)
        assert abs(float(E) - row[4]) < 1e-9 and abs(float(N) - row[5]) < 2e-9
        lon, lat = utm_mp.utm_inverse_mp(row[4], row[5], int(row[2]), bool(row[3]), order=8, dps=30)
        assert abs(float(lon) - row[0]) < 1e-12 and abs(float(lat) - row[1]) < 1e-12
