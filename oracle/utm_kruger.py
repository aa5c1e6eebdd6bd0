"""WGS84 <-> UTM, Karney/Krueger 6th-order series (oracle; PARITY UNPINNED).

Stands in for ``pyproj.Proj("+proj=utm +zone=Z[ +south] +ellps=WGS84 ...")`` as the
reference uses it at EKFGPSSLAM.py:266-271 (forward, called as ``proj(lon, lat)``)
and EKFGPSSLAM.py:291-296 (inverse, ``proj(E, N, inverse=True)``).  pyproj/PROJ is a
third-party dependency that is neither vendored nor version-pinned by the
reference and is not installed in this image, so there is nothing to run it
against: the algorithm below is the published one (C. F. F. Karney,
"Transverse Mercator with an accuracy of a few nanometers", J. Geodesy 2011,
eqs. 7-36; the series PROJ documents for its exact tmerc), truncated at n^6
(error < 1 nm inside a UTM zone).

Also restates the loader-side bookkeeping around the projection:
  * validity mask                      EKFGPSSLAM.py:259
  * zone / hemisphere from the means   EKFGPSSLAM.py:127-134
"""
from __future__ import annotations

import numpy as np

WGS84_A = 6378137.0
WGS84_F = 1.0 / 298.257223563
UTM_K0 = 0.9996
UTM_FE = 500000.0
UTM_FN_SOUTH = 10000000.0


def _series_constants():
    f = WGS84_F
    n = f / (2.0 - f)
    n2, n3, n4, n5, n6 = n * n, n ** 3, n ** 4, n ** 5, n ** 6
    A = WGS84_A / (1.0 + n) * (1.0 + n2 / 4.0 + n4 / 64.0 + n6 / 256.0)
    alpha = np.array([
        n / 2 - 2 * n2 / 3 + 5 * n3 / 16 + 41 * n4 / 180 - 127 * n5 / 288 + 7891 * n6 / 37800,
        13 * n2 / 48 - 3 * n3 / 5 + 557 * n4 / 1440 + 281 * n5 / 630 - 1983433 * n6 / 1935360,
        61 * n3 / 240 - 103 * n4 / 140 + 15061 * n5 / 26880 + 167603 * n6 / 181440,
        49561 * n4 / 161280 - 179 * n5 / 168 + 6601661 * n6 / 7257600,
        34729 * n5 / 80640 - 3418889 * n6 / 1995840,
        212378941 * n6 / 319334400,
    ])
    beta = np.array([
        n / 2 - 2 * n2 / 3 + 37 * n3 / 96 - n4 / 360 - 81 * n5 / 512 + 96199 * n6 / 604800,
        n2 / 48 + n3 / 15 - 437 * n4 / 1440 + 46 * n5 / 105 - 1118711 * n6 / 3870720,
        17 * n3 / 480 - 37 * n4 / 840 - 209 * n5 / 4480 + 5569 * n6 / 90720,
        4397 * n4 / 161280 - 11 * n5 / 504 - 830251 * n6 / 7257600,
        4583 * n5 / 161280 - 108847 * n6 / 3991680,
        20648693 * n6 / 638668800,
    ])
    e2 = f * (2.0 - f)
    return A, alpha, beta, e2


_A, _ALPHA, _BETA, _E2 = _series_constants()
_E = np.sqrt(_E2)


def utm_zone_from_means(lons: np.ndarray, lats: np.ndarray) -> tuple[int, bool]:
    """zone = int((mean(lon)+180)//6+1); south iff mean(lat) < 0  (EKFGPSSLAM.py:127-134)."""
    if lons.size == 0 or lats.size == 0:
        raise ValueError("empty lon/lat")
    zone = int((np.mean(lons) + 180) // 6 + 1)
    south = bool(np.mean(lats) < 0)
    return zone, south


def gnss_validity_mask(lats: np.ndarray, lons: np.ndarray) -> np.ndarray:
    """EKFGPSSLAM.py:259."""
    return (np.abs(lats) <= 90) & (np.abs(lons) <= 180) & (lats != 0) & (lons != 0)


def utm_forward(lon_deg, lat_deg, zone: int, south: bool):
    """(lon, lat) degrees -> (easting, northing) metres in the given UTM zone."""
    lon = np.asarray(lon_deg, dtype=np.float64)
    lat = np.asarray(lat_deg, dtype=np.float64)
    lon0 = np.deg2rad(6.0 * zone - 183.0)
    lam = np.deg2rad(lon) - lon0
    phi = np.deg2rad(lat)
    tau = np.tan(phi)
    sigma = np.sinh(_E * np.arctanh(_E * tau / np.sqrt(1.0 + tau * tau)))
    taup = tau * np.sqrt(1.0 + sigma * sigma) - sigma * np.sqrt(1.0 + tau * tau)
    coslam = np.cos(lam)
    xip = np.arctan2(taup, coslam)
    etap = np.arcsinh(np.sin(lam) / np.sqrt(taup * taup + coslam * coslam))
    xi = xip.copy()
    eta = etap.copy()
    for j in range(1, 7):
        xi = xi + _ALPHA[j - 1] * np.sin(2 * j * xip) * np.cosh(2 * j * etap)
        eta = eta + _ALPHA[j - 1] * np.cos(2 * j * xip) * np.sinh(2 * j * etap)
    easting = UTM_FE + UTM_K0 * _A * eta
    northing = (UTM_FN_SOUTH if south else 0.0) + UTM_K0 * _A * xi
    return easting, northing


def utm_inverse(easting, northing, zone: int, south: bool):
    """(easting, northing) metres -> (lon, lat) degrees."""
    E = np.asarray(easting, dtype=np.float64)
    N = np.asarray(northing, dtype=np.float64)
    lon0 = np.deg2rad(6.0 * zone - 183.0)
    xi = (N - (UTM_FN_SOUTH if south else 0.0)) / (UTM_K0 * _A)
    eta = (E - UTM_FE) / (UTM_K0 * _A)
    xip = xi.copy()
    etap = eta.copy()
    for j in range(1, 7):
        xip = xip - _BETA[j - 1] * np.sin(2 * j * xi) * np.cosh(2 * j * eta)
        etap = etap - _BETA[j - 1] * np.cos(2 * j * xi) * np.sinh(2 * j * eta)
    sh = np.sinh(etap)
    cx = np.cos(xip)
    taup = np.sin(xip) / np.sqrt(sh * sh + cx * cx)
    lam = np.arctan2(sh, cx)
    # Newton on tau'(tau) = taup (Karney eqs. 19-21); converges in 2-3 steps.
    tau = taup / (1.0 - _E2)
    for _ in range(5):
        s1 = np.sqrt(1.0 + tau * tau)
        sigma = np.sinh(_E * np.arctanh(_E * tau / s1))
        taui = tau * np.sqrt(1.0 + sigma * sigma) - sigma * s1
        dtau = (taup - taui) / np.sqrt(1.0 + taui * taui) * (1.0 + (1.0 - _E2) * tau * tau) / ((1.0 - _E2) * s1)
        tau = tau + dtau
    lat = np.rad2deg(np.arctan(tau))
    lon = np.rad2deg(lam + lon0)
    return lon, lat


class KruegerProj:
    """Call-compatible stand-in for ``pyproj.Proj`` built from the proj string the
    reference composes at EKFGPSSLAM.py:267 (only ``+zone=`` and ``+south`` are read)."""

    def __init__(self, proj_string: str):
        zone = None
        for tok in proj_string.split():
            if tok.startswith("+zone="):
                zone = int(tok.split("=")[1])
        if zone is None or not (1 <= zone <= 60):
            raise ValueError(f"bad UTM zone in {proj_string!r}")
        self.zone = zone
        self.south = "+south" in proj_string
        self.srs = proj_string

    def __call__(self, x, y, inverse: bool = False):
        if inverse:
            return utm_inverse(x, y, self.zone, self.south)
        return utm_forward(x, y, self.zone, self.south)
