// Stand-alone kernels behind the reference's per-function surface (one call = one reference
// function, batched over trajectories):
//   K1  utm_forward / utm_inverse / geo_mean       EKFGPSSLAM.py:127-134, :266-271, :291-296
//   K1b associate_spline                            EKFGPSSLAM.py:351-380 (interp1d cubic/linear)
//   K2  sim3 tile statistics + finalize (Umeyama)   EKFGPSSLAM.py:428-459
//   K2b sim3_apply                                  EKFGPSSLAM.py:461-467
//   (K4, the nearest-neighbour error statistics of EKFGPSSLAM.py:1021-1033, lives in gsf_ate.cu)
#include "gsf_common.cuh"
#include "gsf_internal.cuh"

namespace gsf {

// ============================================================================= K1: UTM
// Karney/Krueger 6th-order series (the algorithm PROJ documents for +proj=utm); coefficients
// for WGS84 are evaluated on the host once (gsf_capi.cu) and passed by value.

// xi + sign * sum c_j sin(2j xi') cosh(2j eta'),  eta + sign * sum c_j cos(2j xi') sinh(2j eta'), evaluated with the
// angle-addition recurrences from (sin 2xi', cos 2xi', sinh 2eta', cosh 2eta').
__device__ __forceinline__ void krueger_series_sc(const double* coef, double xi_in, double eta_in, double sign,
                                                  double s2, double c2, double sh2, double ch2, double& xi, double& eta) {
    double sj = s2, cj = c2, shj = sh2, chj = ch2;
    double ax = 0.0, ay = 0.0;
#pragma unroll
    for (int j = 0; j < 6; ++j) {
        ax += coef[j] * sj * chj;
        ay += coef[j] * cj * shj;
        const double sn = sj * c2 + cj * s2, cn = cj * c2 - sj * s2;
        const double shn = shj * ch2 + chj * sh2, chn = chj * ch2 + shj * sh2;
        sj = sn; cj = cn; shj = shn; chj = chn;
    }
    xi = xi_in + sign * ax;
    eta = eta_in + sign * ay;
}
__device__ __forceinline__ void krueger_series(const double* coef, double xi_in, double eta_in, double sign,
                                               double& xi, double& eta) {
    double s2, c2;
    sincos(2.0 * xi_in, &s2, &c2);
    const double e2p = exp(2.0 * eta_in), e2m = 1.0 / e2p;
    krueger_series_sc(coef, xi_in, eta_in, sign, s2, c2, 0.5 * (e2p - e2m), 0.5 * (e2p + e2m), xi, eta);
}

// forward projection of one point (lam = lon - lon0, phi in radians) -> (xi, eta) scaled later by A k0.
// sigma = sinh(e atanh(e sin phi)) has a tiny argument (e sin phi <= 0.082), so both functions are short odd series
// (truncation < 1e-17 relative) and sqrt(1 + sigma^2) is four terms; the double-angle terms of the Krueger series follow
// algebraically from tau', cos lam and sin lam (no sincos(2 xi') / exp(2 eta')).  Inside a zone (|lam| < 0.12 rad = 6.9 deg)
// sin lam, cos lam and asinh(sinh eta') are Taylor series too (truncation < 1e-18): two library calls per point remain
// (sincos phi, atan2); farther from the central meridian the library functions take over.
__device__ __forceinline__ void utm_forward_point(const UtmConst& K, double lam, double phi, double& xi, double& eta) {
    double sp, cp;
    sincos(phi, &sp, &cp);
    const double rcp_c = 1.0 / cp;
    const double tau = sp * rcp_c, t1 = fabs(rcp_c);                    // tan phi, sqrt(1 + tan^2 phi)
    const double x = K.e * sp, x2 = x * x;                             // e tau / t1 = e sin phi
    // atanh(x) = x (1 + x^2/3 + x^4/5 + ... + x^16/17)
    double at = 1.0 / 17.0;
    at = fma(at, x2, 1.0 / 15.0); at = fma(at, x2, 1.0 / 13.0); at = fma(at, x2, 1.0 / 11.0); at = fma(at, x2, 1.0 / 9.0);
    at = fma(at, x2, 1.0 / 7.0); at = fma(at, x2, 1.0 / 5.0); at = fma(at, x2, 1.0 / 3.0); at = fma(at, x2, 1.0);
    const double y = K.e * x * at, y2 = y * y;                         // |y| <= 0.0068
    // sinh(y) = y (1 + y^2/6 + y^4/120 + y^6/5040 + y^8/362880)
    double sg = 1.0 / 362880.0;
    sg = fma(sg, y2, 1.0 / 5040.0); sg = fma(sg, y2, 1.0 / 120.0); sg = fma(sg, y2, 1.0 / 6.0); sg = fma(sg, y2, 1.0);
    const double sigma = y * sg, sg2 = sigma * sigma;                  // sigma^2 <= 4.6e-5
    // sqrt(1 + u) = 1 + u/2 - u^2/8 + u^3/16 - 5 u^4/128 (next term 6e-24)
    const double rt = fma(fma(fma(fma(-5.0 / 128.0, sg2, 1.0 / 16.0), sg2, -0.125), sg2, 0.5), sg2, 1.0);
    const double taup = tau * rt - sigma * t1;
    double sl, cl, etap, v, rh2;
    if (fabs(lam) < 0.12) {
        const double l2 = lam * lam;
        double ps = 1.0 / 6227020800.0;                                // sin: lam (1 - l2/3! + ... + l2^6/13!)
        ps = fma(ps, l2, -1.0 / 39916800.0); ps = fma(ps, l2, 1.0 / 362880.0); ps = fma(ps, l2, -1.0 / 5040.0);
        ps = fma(ps, l2, 1.0 / 120.0); ps = fma(ps, l2, -1.0 / 6.0); ps = fma(ps, l2, 1.0);
        sl = lam * ps;
        double pc = 1.0 / 479001600.0;                                 // cos: 1 - l2/2! + ... + l2^6/12!
        pc = fma(pc, l2, -1.0 / 3628800.0); pc = fma(pc, l2, 1.0 / 40320.0); pc = fma(pc, l2, -1.0 / 720.0);
        pc = fma(pc, l2, 1.0 / 24.0); pc = fma(pc, l2, -0.5); pc = fma(pc, l2, 1.0);
        cl = pc;
        const double rh = rsqrt(taup * taup + cl * cl);
        rh2 = rh * rh;
        v = sl * rh;                                                    // sinh(eta'), |v| <= tan 0.12 = 0.1206
        const double w = v * v;
        // asinh(v) = v (1 - w/6 + 3 w^2/40 - 5 w^3/112 + 35 w^4/1152 - 63 w^5/2816 + 231 w^6/13312 - 143 w^7/10240 + 6435 w^8/557056 - ...)
        double pa = -12155.0 / 1245184.0;
        pa = fma(pa, w, 6435.0 / 557056.0); pa = fma(pa, w, -143.0 / 10240.0); pa = fma(pa, w, 231.0 / 13312.0); pa = fma(pa, w, -63.0 / 2816.0);
        pa = fma(pa, w, 35.0 / 1152.0); pa = fma(pa, w, -5.0 / 112.0); pa = fma(pa, w, 3.0 / 40.0); pa = fma(pa, w, -1.0 / 6.0); pa = fma(pa, w, 1.0);
        etap = v * pa;
    } else {
        sincos(lam, &sl, &cl);
        const double h2 = taup * taup + cl * cl;
        rh2 = 1.0 / h2;
        v = sl * sqrt(rh2);
        etap = asinh(v);
    }
    const double xip = atan2(taup, cl);
    const double s2 = 2.0 * taup * cl * rh2, c2 = (cl * cl - taup * taup) * rh2;
    const double sh2 = 2.0 * v * sqrt(1.0 + v * v), ch2 = 1.0 + 2.0 * v * v;
    krueger_series_sc(K.alpha, xip, etap, 1.0, s2, c2, sh2, ch2, xi, eta);
}

__global__ void utm_forward_kernel(const double* __restrict__ lon, const double* __restrict__ lat, long long n,
                                   const UtmConst K, double* __restrict__ east, double* __restrict__ north) {
    const double D2R = 0.017453292519943295769236907684886;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        double xi, eta;
        utm_forward_point(K, lon[i] * D2R - K.lon0, lat[i] * D2R, xi, eta);
        east[i] = 500000.0 + K.A_k0 * eta;
        north[i] = K.fn + K.A_k0 * xi;
    }
}

// ---- fused GNSS-row ingest (load_gps_data :258-271 + auto_utm_projection :127-134): rows [n,4] = ts, lat, lon, alt
// in the loader's column order.  Row validity (:259): |lat| <= 90, |lon| <= 180, lat != 0, lon != 0.
__device__ __forceinline__ bool gnss_row_valid(double lat, double lon) {
    return fabs(lat) <= 90.0 && fabs(lon) <= 180.0 && lat != 0.0 && lon != 0.0;
}
// Zone guess from the first rows (one warp): out[2] = zone, out[3] = south flag of the mean over the valid rows among the
// first 4096.  The projection runs with this guess while it accumulates the sums of ALL rows; a track whose overall mean
// falls into another zone (or hemisphere) than its beginning is projected again (gnss_rows_zone_kernel raises the flag).
__global__ void gnss_rows_guess_kernel(const double* __restrict__ rows, long long n, double* out) {
    const double2* __restrict__ r2 = reinterpret_cast<const double2*>(rows);
    double a = 0.0, b = 0.0, c = 0.0;
    for (long long i = threadIdx.x; i < n && i < 4096; i += 32) {
        const double2 p = __ldg(r2 + 2 * i), q = __ldg(r2 + 2 * i + 1);
        if (gnss_row_valid(p.y, q.x)) { a += q.x; b += p.y; c += 1.0; }
    }
    a = warp_sum(a); b = warp_sum(b); c = warp_sum(c);
    if (threadIdx.x == 0) {
        out[2] = c > 0.0 ? floor((a / c + 180.0) / 6.0) + 1.0 : 31.0;
        out[3] = (c > 0.0 && b / c < 0.0) ? 1.0 : 0.0;
    }
}
// Stage 2 (one thread): out = mean lon, mean lat, zone, south flag, valid count; part[0] = 1 if the zone / hemisphere
// differ from the guess that the projection used (out[2], out[3] on entry).
__global__ void gnss_rows_zone_kernel(double* part, int nparts, double* out) {
    double a = 0.0, b = 0.0, c = 0.0;
    for (int i = 0; i < nparts; ++i) { a += part[3 * i]; b += part[3 * i + 1]; c += part[3 * i + 2]; }
    a /= c; b /= c;
    const double zone = floor((a + 180.0) / 6.0) + 1.0;           // int((mean_lon + 180) // 6 + 1)
    const double south = (b < 0.0) ? 1.0 : 0.0;
    part[0] = (zone == out[2] && south == out[3]) ? 0.0 : 1.0;
    out[0] = a; out[1] = b; out[2] = zone; out[3] = south; out[4] = c;
}
// Projection: one thread per row, two 128-bit loads; zone / hemisphere are read from device memory (no host
// round trip).  Invalid rows become NaN measurements ("no GNSS at this stamp").  With `part` it also leaves the partial
// sums of the masked means (stage 1 of the fixed-order reduction: fixed grid, fixed stride order -> part[grid][3] = sum lon,
// sum lat, count); with `redo` it runs only if *redo != 0.
__global__ void __launch_bounds__(256) gnss_rows_project_kernel(const double* __restrict__ rows, long long n, UtmConst K,
                                                                const double* __restrict__ zone_dev, double* __restrict__ out_ts,
                                                                double* __restrict__ out_xyz, double* __restrict__ part,
                                                                const double* __restrict__ redo) {
    __shared__ double scratch[3 * 8];
    if (redo && *redo == 0.0) return;
    const double D2R = 0.017453292519943295769236907684886;
    const double lon0 = (6.0 * zone_dev[2] - 183.0) * D2R, fn = zone_dev[3] != 0.0 ? 10000000.0 : 0.0;
    const double2* __restrict__ r2 = reinterpret_cast<const double2*>(rows);
    double v[3] = {0.0, 0.0, 0.0};
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const double2 a = __ldg(r2 + 2 * i), b = __ldg(r2 + 2 * i + 1);       // (ts, lat), (lon, alt)
        double e = nan(""), nn = nan(""), up = nan("");
        if (gnss_row_valid(a.y, b.x)) {
            v[0] += b.x; v[1] += a.y; v[2] += 1.0;
            double xi, eta;
            utm_forward_point(K, b.x * D2R - lon0, a.y * D2R, xi, eta);
            e = 500000.0 + K.A_k0 * eta; nn = fn + K.A_k0 * xi; up = b.y;
        }
        if (out_ts) out_ts[i] = a.x;
        out_xyz[3 * i] = e; out_xyz[3 * i + 1] = nn; out_xyz[3 * i + 2] = up;
    }
    if (part) {
        block_sum<3>(v, scratch);
        if (threadIdx.x == 0) { part[3 * blockIdx.x] = v[0]; part[3 * blockIdx.x + 1] = v[1]; part[3 * blockIdx.x + 2] = v[2]; }
    }
}

__global__ void utm_inverse_kernel(const double* __restrict__ east, const double* __restrict__ north, long long n,
                                   const UtmConst K, double* __restrict__ lon, double* __restrict__ lat) {
    const double R2D = 57.295779513082320876798154814105;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const double xi = (north[i] - K.fn) / K.A_k0, eta = (east[i] - 500000.0) / K.A_k0;
        double xip, etap;
        krueger_series(K.beta, xi, eta, -1.0, xip, etap);
        const double sh = sinh(etap);
        double sx, cx;
        sincos(xip, &sx, &cx);
        const double taup = sx / sqrt(sh * sh + cx * cx);
        const double lam = atan2(sh, cx);
        double tau = taup / (1.0 - K.e2);
        for (int it = 0; it < 5; ++it) {                      // Newton on tau'(tau) = taup
            const double s1 = sqrt(1.0 + tau * tau);
            const double sigma = sinh(K.e * atanh(K.e * tau / s1));
            const double taui = tau * sqrt(1.0 + sigma * sigma) - sigma * s1;
            const double dtau = (taup - taui) / sqrt(1.0 + taui * taui) * (1.0 + (1.0 - K.e2) * tau * tau) / ((1.0 - K.e2) * s1);
            tau += dtau;
        }
        lat[i] = atan(tau) * R2D;
        lon[i] = (lam + K.lon0) * R2D;
    }
}

// Deterministic two-stage sum of (lon, lat) for the zone / hemisphere choice (:131-133).
// Stage 1: fixed grid, fixed per-thread stride order, block partials -> part[grid][2].
__global__ void geo_sum_kernel(const double* __restrict__ lon, const double* __restrict__ lat, long long n, double* part) {
    __shared__ double scratch[2 * 8];
    double v[2] = {0.0, 0.0};
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        v[0] += lon[i]; v[1] += lat[i];
    }
    block_sum<2>(v, scratch);
    if (threadIdx.x == 0) { part[2 * blockIdx.x] = v[0]; part[2 * blockIdx.x + 1] = v[1]; }
}
// Stage 2 (one thread): out[0] = mean lon, out[1] = mean lat, out[2] = zone, out[3] = south flag.
__global__ void geo_finish_kernel(const double* part, int nparts, long long n, double* out) {
    double a = 0.0, b = 0.0;
    for (int i = 0; i < nparts; ++i) { a += part[2 * i]; b += part[2 * i + 1]; }
    a /= (double)n; b /= (double)n;
    out[0] = a; out[1] = b;
    out[2] = floor((a + 180.0) / 6.0) + 1.0;           // int((mean_lon + 180) // 6 + 1)
    out[3] = (b < 0.0) ? 1.0 : 0.0;
}

// ============================================================================= K1b: spline association
// One block per trajectory.  GNSS stamps must be sorted and unique (the host does the argsort /
// np.unique of :340-349).  Segments are split at gaps > threshold (:351-354); per segment a
// not-a-knot cubic spline (>= 4 knots, scipy make_interp_spline(k=3) default boundary
// conditions) or a linear interpolant (2-3 knots) is evaluated at the SLAM stamps inside
// [seg start, seg end]; everything else is NaN / invalid (:372-379).
// The spline is solved in moment form (second derivatives m_j) with the not-a-knot rows
// eliminated, by the Thomas algorithm: thread a in {0,1,2} owns coordinate axis a.
// work: per GNSS sample 4 doubles (c', and 3 moments) -> caller-provided workspace.

__device__ void spline_moments(const double* t, const double* y /*stride 3*/, int m, double* cp, double* mom /*stride 3*/, int axis) {
    // interior unknowns m_1..m_{m-2}; not-a-knot: m_0 = ((h0+h1) m_1 - h0 m_2)/h1, same at the end.
    const int nu = m - 2;
    auto h = [&](int j) { return t[j + 1] - t[j]; };
    auto rhs = [&](int j) {   // row for knot j (1..m-2)
        return 6.0 * ((y[3 * (j + 1) + axis] - y[3 * j + axis]) / h(j) - (y[3 * j + axis] - y[3 * (j - 1) + axis]) / h(j - 1));
    };
    // row j: lo*m_{j-1} + di*m_j + up*m_{j+1} = rhs, with end substitutions folded in.
    double prev_cp = 0.0, prev_dp = 0.0;
    for (int j = 1; j <= nu; ++j) {
        double lo = h(j - 1), di = 2.0 * (h(j - 1) + h(j)), up = h(j);
        if (j == 1) {          // substitute m_0
            const double h0 = h(0), h1 = h(1);
            di += h0 * (h0 + h1) / h1; up -= h0 * h0 / h1; lo = 0.0;
        }
        if (j == nu) {         // substitute m_{m-1}
            const double ha = h(m - 3), hb = h(m - 2);
            di += hb * (ha + hb) / ha; lo -= hb * hb / ha; up = 0.0;
        }
        if (nu == 1) { lo = 0.0; up = 0.0; }
        const double den = di - lo * prev_cp;
        const double c = up / den;
        const double d = (rhs(j) - lo * prev_dp) / den;
        cp[j] = c; mom[3 * j + axis] = d;
        prev_cp = c; prev_dp = d;
    }
    for (int j = nu - 1; j >= 1; --j) mom[3 * j + axis] -= cp[j] * mom[3 * (j + 1) + axis];
    {
        const double h0 = h(0), h1 = h(1);
        mom[axis] = ((h0 + h1) * mom[3 + axis] - h0 * mom[6 + axis]) / h1;
        const double ha = h(m - 3), hb = h(m - 2);
        mom[3 * (m - 1) + axis] = ((ha + hb) * mom[3 * (m - 2) + axis] - hb * mom[3 * (m - 3) + axis]) / ha;
    }
}

__global__ void associate_kernel(const AssocArgs A) {
    __shared__ int seg_lo, seg_hi, seg_ok, more;
    for (int b = blockIdx.x; b < A.B; b += gridDim.x) {
        const long long g0 = A.gps_off[b], s0 = A.slam_off[b];
        const int M = (int)(A.gps_off[b + 1] - g0), N = (int)(A.slam_off[b + 1] - s0);
        const double* gt = A.gps_t + g0; const double* gy = A.gps_xyz + 3 * g0;
        const double* st = A.slam_t + s0;
        double* out = A.aligned + 3 * s0; unsigned char* val = A.valid + s0;
        double* cp = A.work + 4 * g0; double* mom = cp + M;      // cp[M], mom[3M]
        for (int i = threadIdx.x; i < N; i += blockDim.x) {
            out[3 * i] = out[3 * i + 1] = out[3 * i + 2] = nan(""); val[i] = 0;
        }
        if (M < 2 || N == 0) { __syncthreads(); continue; }
        int lo = 0;
        while (true) {
            __syncthreads();
            if (threadIdx.x == 0) {
                int hi = lo; bool ok = true;
                while (hi + 1 < M && !(gt[hi + 1] - gt[hi] > A.gap)) { if (!(gt[hi + 1] - gt[hi] > 1e-9)) ok = false; ++hi; }
                seg_lo = lo; seg_hi = hi; seg_ok = ok && (hi - lo + 1 >= 2); more = (hi + 1 < M);
            }
            __syncthreads();
            const int a0 = seg_lo, a1 = seg_hi, m = a1 - a0 + 1;
            if (seg_ok) {
                if (m >= 4 && threadIdx.x < 3) spline_moments(gt + a0, gy + 3 * a0, m, cp + a0, mom + 3 * a0, threadIdx.x);
                __syncthreads();
                const double ta = gt[a0], tb = gt[a1];
                for (int i = threadIdx.x; i < N; i += blockDim.x) {
                    const double t = st[i];
                    if (!(t >= ta && t <= tb)) continue;                 // scipy: NaN outside [ta, tb]
                    int l = a0, r = a1;                                   // largest j with gt[j] <= t, j < a1
                    while (r - l > 1) { int mid = (l + r) >> 1; if (gt[mid] <= t) l = mid; else r = mid; }
                    const int j = l;
                    const double hh = gt[j + 1] - gt[j];
                    for (int ax = 0; ax < 3; ++ax) {
                        const double y0 = gy[3 * j + ax], y1 = gy[3 * (j + 1) + ax];
                        double v;
                        if (m >= 4) {
                            const double wa = (gt[j + 1] - t) / hh, wb = (t - gt[j]) / hh;
                            const double m0 = mom[3 * j + ax], m1 = mom[3 * (j + 1) + ax];
                            v = wa * y0 + wb * y1 + ((wa * wa * wa - wa) * m0 + (wb * wb * wb - wb) * m1) * (hh * hh) / 6.0;
                        } else {
                            v = (y1 - y0) / hh * (t - gt[j]) + y0;        // interp1d kind='linear'
                        }
                        out[3 * i + ax] = v;
                    }
                    val[i] = !row_has_nan(out[3 * i], out[3 * i + 1], out[3 * i + 2]);
                }
            }
            if (!more) break;
            lo = seg_hi + 1;
        }
        __syncthreads();
    }
}

// ============================================================================= K2: Umeyama (stand-alone)
// Tile statistics: every block reduces up to TILE points of one trajectory to
// (n, mean_src, mean_dst, H about the tile means, ss) with a tile-local pivot; the finalize
// kernel merges tiles in order with the pairwise-covariance update, then SVD.  One pass over
// the data, any trajectory length, bit-reproducible.
constexpr int SIM3_TILE = 8192;
constexpr int SIM3_STAT = 20;    // n, mu_s[3], mu_d[3], H[9], ss, pad[3]
// Tile length as a function of the longest trajectory only (so results do not depend on the device): 8192
// points up to 8M-point trajectories, beyond that at most 1024 tiles per trajectory.
__host__ __device__ inline long long sim3_tile_len(long long max_len) {
    if (max_len <= (long long)SIM3_TILE * 1024) return SIM3_TILE;
    const long long t = (max_len + 1023) / 1024;
    return (t + 511) / 512 * 512;
}

__device__ __forceinline__ void sim3_accumulate(double* v, const double* piv, double s0, double s1, double s2, double d0, double d1, double d2) {
    const double a0 = s0 - piv[0], a1 = s1 - piv[1], a2 = s2 - piv[2];
    const double b0 = d0 - piv[3], b1 = d1 - piv[4], b2 = d2 - piv[5];
    v[0] += 1.0; v[1] += a0; v[2] += a1; v[3] += a2; v[4] += b0; v[5] += b1; v[6] += b2;
    v[7] += a0 * b0; v[8] += a0 * b1; v[9] += a0 * b2;
    v[10] += a1 * b0; v[11] += a1 * b1; v[12] += a1 * b2;
    v[13] += a2 * b0; v[14] += a2 * b1; v[15] += a2 * b2;
    v[16] += a0 * a0 + a1 * a1 + a2 * a2;
}

__global__ void __launch_bounds__(256) sim3_tile_stats_kernel(const double* __restrict__ src, const double* __restrict__ dst,
                                                              const long long* __restrict__ offsets,
                                                              const unsigned char* __restrict__ mask, int tiles_max, long long tile_len,
                                                              double* __restrict__ stats) {
    __shared__ double scratch[17 * 8];
    __shared__ double piv[6];
    __shared__ int piv_idx;
    const int b = blockIdx.y, tile = blockIdx.x;
    const long long e0 = offsets[b];
    const long long n = offsets[b + 1] - e0;
    double* o = stats + ((size_t)b * tiles_max + tile) * SIM3_STAT;
    const long long lo = (long long)tile * tile_len;
    if (lo >= n) { if (threadIdx.x == 0) o[0] = 0.0; return; }
    const long long hi = min(n, lo + tile_len);
    if (threadIdx.x == 0) piv_idx = 0x7fffffff;
    __syncthreads();
    for (long long i = lo + threadIdx.x; i < hi; i += blockDim.x)
        if (!mask || mask[e0 + i]) { atomicMin(&piv_idx, (int)(i - lo)); break; }
    __syncthreads();
    if (piv_idx == 0x7fffffff) { if (threadIdx.x == 0) o[0] = 0.0; return; }
    if (threadIdx.x < 6) {
        const long long g = e0 + lo + piv_idx;
        piv[threadIdx.x] = threadIdx.x < 3 ? src[3 * g + threadIdx.x] : dst[3 * g + threadIdx.x - 3];
    }
    __syncthreads();
    double v[17];
#pragma unroll
    for (int k = 0; k < 17; ++k) v[k] = 0.0;
    // Point pairs (2p, 2p+1) of the tile, pairs strided over the threads: the same summation order whether the
    // rows are 16-byte aligned (three 128-bit loads per array and pair) or not (64-bit loads).
    const long long g0 = e0 + lo, cnt = hi - lo, npairs = (cnt + 1) >> 1;
    const bool vec = !(g0 & 1) && ((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(dst)) & 15) == 0;
    const double2* __restrict__ s2 = reinterpret_cast<const double2*>(src + 3 * g0);
    const double2* __restrict__ d2 = reinterpret_cast<const double2*>(dst + 3 * g0);
    for (long long p = threadIdx.x; p < npairs; p += 256) {
        const long long i = 2 * p;
        const bool two = i + 1 < cnt;
        const bool m0 = !mask || mask[g0 + i], m1 = two && (!mask || mask[g0 + i + 1]);
        if (vec && two) {
            const double2 a0 = __ldg(s2 + 3 * p), a1 = __ldg(s2 + 3 * p + 1), a2 = __ldg(s2 + 3 * p + 2);
            const double2 b0 = __ldg(d2 + 3 * p), b1 = __ldg(d2 + 3 * p + 1), b2 = __ldg(d2 + 3 * p + 2);
            if (m0) sim3_accumulate(v, piv, a0.x, a0.y, a1.x, b0.x, b0.y, b1.x);
            if (m1) sim3_accumulate(v, piv, a1.y, a2.x, a2.y, b1.y, b2.x, b2.y);
        } else {
            const long long g = g0 + i;
            if (m0) sim3_accumulate(v, piv, src[3 * g], src[3 * g + 1], src[3 * g + 2], dst[3 * g], dst[3 * g + 1], dst[3 * g + 2]);
            if (m1) sim3_accumulate(v, piv, src[3 * g + 3], src[3 * g + 4], src[3 * g + 5], dst[3 * g + 3], dst[3 * g + 4], dst[3 * g + 5]);
        }
    }
    block_sum<17>(v, scratch);
    if (threadIdx.x == 0) {
        const double cnt_v = v[0], inv = 1.0 / cnt_v;
        const double ma[3] = {v[1] * inv, v[2] * inv, v[3] * inv}, mb[3] = {v[4] * inv, v[5] * inv, v[6] * inv};
        o[0] = cnt_v;
        for (int k = 0; k < 3; ++k) { o[1 + k] = piv[k] + ma[k]; o[4 + k] = piv[3 + k] + mb[k]; }
        for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) o[7 + 3 * r + c] = v[7 + 3 * r + c] - cnt_v * ma[r] * mb[c];
        o[16] = v[16] - cnt_v * (ma[0] * ma[0] + ma[1] * ma[1] + ma[2] * ma[2]);
    }
}

// Pairwise-covariance merge of two (n, means, centred H, ss) records, `a` before `b` in point order.
__device__ __forceinline__ void sim3_merge(double* a, const double* b) {
    const double m = b[0];
    if (m == 0.0) return;
    if (a[0] == 0.0) {
#pragma unroll
        for (int k = 0; k < 17; ++k) a[k] = b[k];
        return;
    }
    const double n = a[0], tot = n + m, w = n * m / tot;
    double da[3], db[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) { da[k] = b[1 + k] - a[1 + k]; db[k] = b[4 + k] - a[4 + k]; }
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int c = 0; c < 3; ++c) a[7 + 3 * r + c] += b[7 + 3 * r + c] + w * da[r] * db[c];
    a[16] += b[16] + w * (da[0] * da[0] + da[1] * da[1] + da[2] * da[2]);
#pragma unroll
    for (int k = 0; k < 3; ++k) { a[1 + k] += da[k] * (m / tot); a[4 + k] += db[k] * (m / tot); }
    a[0] = tot;
}

// One warp per trajectory: every lane merges a contiguous run of tiles in tile order, then the lanes are merged
// by a fixed in-order tree (lane L absorbs lane L + o for o = 1, 2, 4, ...) -> deterministic; lane 0 finishes.
// stats_out != NULL: the merged statistics (SIM3_STAT doubles per trajectory) are stored and nothing else -- one shard's
// contribution to a trajectory that is spread over several GPUs (gsf_sim3_partial_stats_dev).
__global__ void __launch_bounds__(128) sim3_finalize_kernel(const double* __restrict__ stats, int tiles_max, int B,
                                                            double* __restrict__ Rout, double* __restrict__ tout, double* __restrict__ sout,
                                                            int* __restrict__ status, double* __restrict__ stats_out) {
    const int b = blockIdx.x * 4 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (b >= B) return;
    double acc[17];
#pragma unroll
    for (int k = 0; k < 17; ++k) acc[k] = 0.0;
    const int per = (tiles_max + 31) / 32;
    for (int tl = lane * per; tl < min(tiles_max, (lane + 1) * per); ++tl) sim3_merge(acc, stats + ((size_t)b * tiles_max + tl) * SIM3_STAT);
    if (tiles_max > 1) {
#pragma unroll 1
        for (int o = 1; o < 32; o <<= 1) {
            double other[17];
#pragma unroll
            for (int k = 0; k < 17; ++k) other[k] = __shfl_down_sync(GSF_FULL_MASK, acc[k], o);
            if ((lane & (2 * o - 1)) == 0) sim3_merge(acc, other);
        }
    }
    if (lane != 0) return;
    if (stats_out) {
#pragma unroll
        for (int k = 0; k < SIM3_STAT; ++k) stats_out[SIM3_STAT * (size_t)b + k] = k < 17 ? acc[k] : 0.0;
        return;
    }
    double* R = Rout + 9 * (size_t)b; double* t = tout + 3 * (size_t)b;
    if (acc[0] < 3.0) {                                 // (None, None, None), :431
        for (int k = 0; k < 9; ++k) R[k] = nan("");
        t[0] = t[1] = t[2] = nan(""); sout[b] = nan(""); status[b] = ST_TOO_FEW_POINTS;
        return;
    }
    double s;
    status[b] = umeyama_finish((int)acc[0], acc + 1, acc + 4, acc + 7, acc[16], R, t, s);
    sout[b] = s;
}

// ============================================================================= K2b: Sim3 apply
// p' = s R p + t ; q' = normalize(q_R (x) normalize(q)).  One thread per pose; positions go
// through shared memory so global accesses stay fully coalesced (24-byte rows).
__global__ void __launch_bounds__(256) sim3_apply_kernel(const double* __restrict__ pos, const double* __restrict__ quat,
                                                         const long long* __restrict__ offsets,
                                                         const double* __restrict__ Rm, const double* __restrict__ tv,
                                                         const double* __restrict__ sv,
                                                         double* __restrict__ out_pos, double* __restrict__ out_quat,
                                                         int* __restrict__ status) {
    __shared__ double tile[256 * 3];
    __shared__ double xf[20];
    const int b = blockIdx.y;
    const long long e0 = offsets[b], n = offsets[b + 1] - e0;
    if (threadIdx.x == 0) {
        Quat qR = quat_from_matrix(Rm + 9 * (size_t)b);
        xf[12] = qR.x; xf[13] = qR.y; xf[14] = qR.z; xf[15] = qR.w;
    }
    if (threadIdx.x < 9) xf[threadIdx.x] = Rm[9 * (size_t)b + threadIdx.x];
    if (threadIdx.x < 3) xf[9 + threadIdx.x] = tv[3 * (size_t)b + threadIdx.x];
    if (threadIdx.x == 9) xf[16] = sv[b];
    __syncthreads();
    const Quat qR{xf[12], xf[13], xf[14], xf[15]};
    const double s = xf[16];
    int bad = 0;
    for (long long lo = (long long)blockIdx.x * 256; lo < n; lo += (long long)gridDim.x * 256) {
        const int cnt = (int)min((long long)256, n - lo);
        __syncthreads();
        for (int k = threadIdx.x; k < 3 * cnt; k += 256) tile[k] = pos[3 * (e0 + lo) + k];
        __syncthreads();
        if (threadIdx.x < cnt) {
            const double x = tile[3 * threadIdx.x], y = tile[3 * threadIdx.x + 1], z = tile[3 * threadIdx.x + 2];
            double rx, ry, rz;
            mat_vec(xf, x, y, z, rx, ry, rz);
            tile[3 * threadIdx.x] = s * rx + xf[9]; tile[3 * threadIdx.x + 1] = s * ry + xf[10]; tile[3 * threadIdx.x + 2] = s * rz + xf[11];
            const long long g = e0 + lo + threadIdx.x;
            const double2 q01 = reinterpret_cast<const double2*>(quat)[2 * g], q23 = reinterpret_cast<const double2*>(quat)[2 * g + 1];
            Quat q{q01.x, q01.y, q23.x, q23.y};
            if (qnorm2(q) == 0.0) bad = 1;
            Quat r = qunit(qmul(qR, qunit(q)));
            reinterpret_cast<double2*>(out_quat)[2 * g] = make_double2(r.x, r.y);
            reinterpret_cast<double2*>(out_quat)[2 * g + 1] = make_double2(r.z, r.w);
        }
        __syncthreads();
        for (int k = threadIdx.x; k < 3 * cnt; k += 256) out_pos[3 * (e0 + lo) + k] = tile[k];
    }
    if (bad) atomicOr(status + b, ST_BAD_QUATERNION);
}

// ----------------------------------------------------------------------------- launchers
cudaError_t launch_utm(bool inverse, const double* a, const double* b, long long n, const UtmConst& K, double* o1, double* o2,
                       int num_sms, cudaStream_t stream) {
    if (n <= 0) return cudaSuccess;
    long long blocks = (n + 255) / 256;
    if (blocks > (long long)num_sms * 16) blocks = (long long)num_sms * 16;
    if (inverse) utm_inverse_kernel<<<(unsigned)blocks, 256, 0, stream>>>(a, b, n, K, o1, o2);
    else utm_forward_kernel<<<(unsigned)blocks, 256, 0, stream>>>(a, b, n, K, o1, o2);
    return cudaGetLastError();
}
cudaError_t launch_gnss_rows(const double* rows, long long n, const UtmConst& K, double* part, int nparts, double* zone_out,
                             double* out_ts, double* out_xyz, int num_sms, cudaStream_t stream) {
    if (n <= 0) return cudaSuccess;
    // one pass over the rows: projection with the zone of the track's beginning + the sums of the masked means; the zone of
    // the whole track is known afterwards, and the (rare) track that needs another zone is projected a second time
    gnss_rows_guess_kernel<<<1, 32, 0, stream>>>(rows, n, zone_out);
    gnss_rows_project_kernel<<<nparts, 256, 0, stream>>>(rows, n, K, zone_out, out_ts, out_xyz, part, nullptr);
    gnss_rows_zone_kernel<<<1, 1, 0, stream>>>(part, nparts, zone_out);
    gnss_rows_project_kernel<<<nparts, 256, 0, stream>>>(rows, n, K, zone_out, out_ts, out_xyz, nullptr, part);
    (void)num_sms;
    return cudaGetLastError();
}
cudaError_t launch_geo_mean(const double* lon, const double* lat, long long n, double* part, int nparts, double* out, cudaStream_t stream) {
    geo_sum_kernel<<<nparts, 256, 0, stream>>>(lon, lat, n, part);
    geo_finish_kernel<<<1, 1, 0, stream>>>(part, nparts, n, out);
    return cudaGetLastError();
}
cudaError_t launch_associate(const AssocArgs& a, int num_sms, cudaStream_t stream) {
    if (a.B <= 0) return cudaSuccess;
    int grid = a.B < num_sms * 8 ? a.B : num_sms * 8;
    associate_kernel<<<grid, 128, 0, stream>>>(a);
    return cudaGetLastError();
}
int sim3_tiles_for(long long max_len) { const long long tl = sim3_tile_len(max_len); const long long t = (max_len + tl - 1) / tl; return t > 0 ? (int)t : 1; }
cudaError_t launch_umeyama(const double* src, const double* dst, const long long* offsets, const unsigned char* mask,
                           int B, long long max_len, double* work, double* R, double* t, double* s, int* status, cudaStream_t stream) {
    if (B <= 0) return cudaSuccess;
    const int tiles = sim3_tiles_for(max_len);
    for (int b0 = 0; b0 < B; b0 += 65535) {                    // gridDim.y limit
        const int nb = (B - b0) < 65535 ? (B - b0) : 65535;
        dim3 grid(tiles, nb);
        sim3_tile_stats_kernel<<<grid, 256, 0, stream>>>(src, dst, offsets + b0, mask, tiles, sim3_tile_len(max_len), work + (size_t)b0 * tiles * SIM3_STAT);
    }
    sim3_finalize_kernel<<<(B + 3) / 4, 128, 0, stream>>>(work, tiles, B, R, t, s, status, nullptr);
    return cudaGetLastError();
}
// One trajectory spread over several GPUs: a shard's statistics (count, means, centred cross-covariance, sum |src_c|^2:
// SIM3_STAT doubles) ...
cudaError_t launch_sim3_partial_stats(const double* src, const double* dst, const long long* offsets2, const unsigned char* mask, long long n,
                                      double* work, double* stats_out, cudaStream_t stream) {
    const int tiles = sim3_tiles_for(n);
    dim3 grid(tiles, 1);
    sim3_tile_stats_kernel<<<grid, 256, 0, stream>>>(src, dst, offsets2, mask, tiles, sim3_tile_len(n), work);
    sim3_finalize_kernel<<<1, 128, 0, stream>>>(work, tiles, 1, nullptr, nullptr, nullptr, nullptr, stats_out);
    return cudaGetLastError();
}
// ... and R, t, s from the statistics of all shards, merged in shard order with the pairwise covariance update.
cudaError_t launch_sim3_from_partial_stats(const double* stats, int shards, double* R, double* t, double* s, int* status, cudaStream_t stream) {
    sim3_finalize_kernel<<<1, 128, 0, stream>>>(stats, shards, 1, R, t, s, status, nullptr);
    return cudaGetLastError();
}
cudaError_t launch_sim3_apply(const double* pos, const double* quat, const long long* offsets, const double* R, const double* t,
                              const double* s, int B, long long max_len, double* out_pos, double* out_quat, int* status,
                              int num_sms, cudaStream_t stream) {
    if (B <= 0) return cudaSuccess;
    long long tiles = (max_len + 255) / 256; if (tiles < 1) tiles = 1;
    long long cap = (long long)num_sms * 32; if (tiles > cap) tiles = cap;
    for (int b0 = 0; b0 < B; b0 += 65535) {
        const int nb = (B - b0) < 65535 ? (B - b0) : 65535;
        dim3 grid((unsigned)tiles, nb);
        sim3_apply_kernel<<<grid, 256, 0, stream>>>(pos, quat, offsets + b0, R + 9 * (size_t)b0, t + 3 * (size_t)b0, s + b0,
                                                    out_pos, out_quat, status + b0);
    }
    return cudaGetLastError();
}
}  // namespace gsf
