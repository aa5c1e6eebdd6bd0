"""Independent high-precision pin of the UTM projection (test infrastructure; never imported by the product).

The reference projects with pyproj/PROJ (EKFGPSSLAM.py:266-271, :291-296), which is absent here, and
``oracle/utm_kruger.py`` restates Karney's 6th-order Krueger series from a recalled coefficient table.
This module does NOT use that table.  It evaluates the *definition* of the ellipsoidal transverse
Mercator projection at 40 significant digits with mpmath:

  * conformal latitude chi(phi) and the Gauss-Schreiber coordinates zeta' = xi' + i eta' on the
    conformal sphere (closed forms);
  * the TM map is the analytic function zeta = f(zeta') that is real on the real axis and equals the
    rectifying latitude mu(chi) there (true scale on the central meridian): f(zeta') = zeta' +
    sum_j a_j sin(2 j zeta'), with a_j the Fourier coefficients of mu(chi) - chi.  The a_j are
    obtained here by *numerical quadrature* of the exact mu(chi) (meridian arc from the incomplete
    elliptic integral of the second kind), to as many orders as asked (default 12), so neither the
    table nor its truncation order enters;
  * the inverse coefficients b_j likewise from chi(mu) - mu.

``python -m oracle.utm_mp`` regenerates tests/golden/utm_mp.npz (lattice of +-3.5 deg x +-84 deg plus
the shipped data's neighbourhoods) and prints the derived coefficients next to the table's.
"""
from __future__ import annotations

import mpmath as mp

A_WGS84 = mp.mpf(6378137)
F_WGS84 = 1 / mp.mpf("298.257223563")
K0 = mp.mpf("0.9996")


def _setup(dps=40):
    mp.mp.dps = dps
    f = F_WGS84
    e2 = f * (2 - f)
    return e2, mp.sqrt(e2)


def chi_of_phi(phi, e):
    """conformal latitude (exact)."""
    return mp.atan(mp.sinh(mp.asinh(mp.tan(phi)) - e * mp.atanh(e * mp.sin(phi))))


def meridian_arc(phi, e2):
    """meridian arc length from the equator on the unit-semi-major-axis ellipsoid:
    (1 - e^2) int_0^phi (1 - e^2 sin^2 t)^(-3/2) dt = E(phi | e^2) - e^2 sin phi cos phi / sqrt(1 - e^2 sin^2 phi)."""
    s, c = mp.sin(phi), mp.cos(phi)
    return mp.ellipe(phi, e2) - e2 * s * c / mp.sqrt(1 - e2 * s * s)


def derive_coefficients(order=12, dps=40):
    """-> (A_over_a, alpha[1..order], beta[1..order]) by quadrature of the exact latitude functions."""
    e2, e = _setup(dps)
    quarter = meridian_arc(mp.pi / 2, e2)            # M(pi/2) / a
    A_over_a = 2 * quarter / mp.pi

    def phi_from_chi(chi):
        if chi == 0:
            return mp.mpf(0)
        return mp.findroot(lambda p: chi_of_phi(p, e) - chi, chi, tol=mp.mpf(10) ** (-dps + 4))

    def mu_of_chi(chi):
        return meridian_arc(phi_from_chi(chi), e2) / A_over_a

    def phi_from_mu(mu):
        if mu == 0:
            return mp.mpf(0)
        return mp.findroot(lambda p: meridian_arc(p, e2) / A_over_a - mu, mu, tol=mp.mpf(10) ** (-dps + 4))

    def chi_of_mu(mu):
        return chi_of_phi(phi_from_mu(mu), e)

    half = mp.pi / 2
    # tabulate f on Gauss-Legendre nodes once (the root solves are the expensive part), reuse for every j
    nodes = mp.calculus.quadrature.GaussLegendre(mp.mp).calc_nodes(6, mp.mp.prec)      # degree 6: 3*2^5 = 96 nodes on [-1, 1]
    xs = [(half / 2) * (x + 1) for x, _ in nodes]
    ws = [(half / 2) * w for _, w in nodes]
    fa = [mu_of_chi(x) - x for x in xs]
    fb = [chi_of_mu(x) - x for x in xs]
    alpha = [4 / mp.pi * sum(w * f * mp.sin(2 * j * x) for x, w, f in zip(xs, ws, fa)) for j in range(1, order + 1)]
    beta = [-4 / mp.pi * sum(w * f * mp.sin(2 * j * x) for x, w, f in zip(xs, ws, fb)) for j in range(1, order + 1)]
    return A_over_a, alpha, beta


_CACHE = {}


def coefficients(order=12, dps=40):
    key = (order, dps)
    if key not in _CACHE:
        _CACHE[key] = derive_coefficients(order, dps)
    return _CACHE[key]


def utm_forward_mp(lon_deg, lat_deg, zone, south, order=12, dps=40):
    """(lon, lat) degrees -> (easting, northing) as mpf, from the derived series."""
    e2, e = _setup(dps)
    A_over_a, alpha, _ = coefficients(order, dps)
    lam = mp.radians(mp.mpf(lon_deg)) - mp.radians(mp.mpf(6 * zone - 183))
    chi = chi_of_phi(mp.radians(mp.mpf(lat_deg)), e)
    xip = mp.atan2(mp.sin(chi), mp.cos(chi) * mp.cos(lam))
    etap = mp.atanh(mp.cos(chi) * mp.sin(lam))
    zp = mp.mpc(xip, etap)
    z = zp + sum(a * mp.sin(2 * (j + 1) * zp) for j, a in enumerate(alpha))
    scale = K0 * A_WGS84 * A_over_a
    return 500000 + scale * z.imag, (10000000 if south else 0) + scale * z.real


def utm_inverse_mp(easting, northing, zone, south, order=12, dps=40):
    """(E, N) metres -> (lon, lat) degrees as mpf."""
    e2, e = _setup(dps)
    A_over_a, _, beta = coefficients(order, dps)
    scale = K0 * A_WGS84 * A_over_a
    z = mp.mpc((mp.mpf(northing) - (10000000 if south else 0)) / scale, (mp.mpf(easting) - 500000) / scale)
    zp = z - sum(b * mp.sin(2 * (j + 1) * z) for j, b in enumerate(beta))
    xip, etap = zp.real, zp.imag
    chi = mp.asin(mp.sin(xip) / mp.cosh(etap))
    lam = mp.atan2(mp.sinh(etap), mp.cos(xip))
    phi = mp.findroot(lambda p: chi_of_phi(p, e) - chi, chi, tol=mp.mpf(10) ** (-dps + 4)) if chi != 0 else mp.mpf(0)
    return mp.degrees(lam) + (6 * zone - 183), mp.degrees(phi)


def lattice():
    """test points: (lon offset from the central meridian, lat) lattice over a widened UTM zone, both hemispheres,
    plus the neighbourhoods of the shipped fixtures (EKFGPSSLAM pair A is projected at lat 8.39 / lon 49.03, zone 39;
    pair B at lat 49.0 / lon 8.4, zone 32)."""
    pts = []
    for dlon in (-3.5, -3.0, -2.0, -1.0, -0.25, 0.0, 0.5, 1.5, 2.5, 3.0, 3.5):
        for lat in (-84.0, -80.0, -65.5, -45.0, -23.4, -8.0, -0.5, 0.0, 1e-6, 0.3, 8.39, 23.4, 35.0, 49.0, 60.1, 72.0, 80.0, 84.0):
            for zone in (1, 32, 39, 60):
                pts.append((6 * zone - 183 + dlon, lat, zone, lat < 0))
    pts += [(49.03 + 0.001 * k, 8.39 + 0.0007 * k, 39, False) for k in range(-5, 6)]
    pts += [(8.4 + 0.002 * k, 49.0 - 0.0013 * k, 32, False) for k in range(-5, 6)]
    return pts


def main():
    import os
    import numpy as np
    from oracle import utm_kruger as uk
    A_over_a, alpha, beta = coefficients()
    print("A / a derived", mp.nstr(A_over_a, 25), " table", repr(uk._A / uk.WGS84_A))
    for j in range(8):
        ta = uk._ALPHA[j] if j < 6 else 0.0
        tb = uk._BETA[j] if j < 6 else 0.0
        print(f"j={j + 1}: alpha derived {mp.nstr(alpha[j], 22)} table {ta!r} | beta derived {mp.nstr(beta[j], 22)} table {tb!r}")
    pts = lattice()
    rows = []
    for lon, lat, zone, south in pts:
        E, N = utm_forward_mp(lon, lat, zone, south)
        rows.append((lon, lat, zone, 1.0 if south else 0.0, float(E), float(N)))
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden", "utm_mp.npz")
    np.savez_compressed(out, points=np.array(rows, dtype=np.float64),
                        alpha=np.array([float(a) for a in alpha]), beta=np.array([float(b) for b in beta]), A_over_a=float(A_over_a))
    print("wrote", os.path.normpath(out), len(rows), "points")


if __name__ == "__main__":
    main()
