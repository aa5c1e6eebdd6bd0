// K1b for ONE long trajectory: dynamic_time_alignment (/root/reference/EKFGPSSLAM.py:325-387) when the GNSS track has
// millions of samples (BASELINE config 4: 1e8 poses).  The per-trajectory kernel of gsf_kernels.cu solves the not-a-knot
// spline of a segment with one serial Thomas sweep per axis -- fine for 271 knots, hopeless for 1e8.  Here the solve is
// LOCAL: the spline's tridiagonal system is diagonally dominant (|off-diagonal| <= diagonal / 2 for any knot spacing), so the
// influence of a boundary value on the moment k knots away decays at least like (2 - sqrt 3)^k = 0.268^k.  Every thread
// owns a chunk of AL_CH = 64 consecutive knots and solves the system on the chunk widened by a halo of AL_H = 20 knots on both
// sides, with a natural end (m = 0) where the halo cuts the segment and the true not-a-knot rows where the segment really
// ends inside the halo; only the chunk's own moments are kept.  The cut moment is wrong by its own size, and 0.268^20 =
// 3.7e-12 of that reaches the chunk: for GNSS noise of 0.3 m at 10 Hz (second differences of 30 m/s^2, a curvature term of
// m h^2 / 8 = 4 cm in the interpolant) that is 1.5e-13 m, far below one ulp of a UTM coordinate (9.3e-10 m) -- the values
// equal the global solve's (the test compares with scipy and with the serial kernel).  (32-knot chunks with a
// 32-knot halo: 12.7 ms for 1e8 knots, three times the arithmetic and the local-memory traffic per knot.)  Segments split at gaps > max_gps_gap_threshold
// (:351-354); 2-3 knot segments are linear (:362), single knots give nothing (:361).
// The evaluation is one thread per SLAM stamp: binary search for its knot interval, segment membership from the gaps
// around it, cubic in moment form / linear / NaN (scipy: NaN outside [seg start, seg end], :377-379).
#include "gsf_common.cuh"
#include "gsf_internal.cuh"

namespace gsf {

constexpr int AL_CH = 64;
constexpr int AL_H = 20;
constexpr int AL_MAX = AL_CH + 2 * AL_H + 2;

__global__ void __launch_bounds__(128) assoc_long_moments_kernel(const double* __restrict__ gt, const double* __restrict__ gy, long long M, double gap,
                                                                 double* __restrict__ mom, int* __restrict__ bad_steps) {
    const long long chunk = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long j0 = chunk * AL_CH;
    if (j0 >= M) return;
    const long long j1 = min(j0 + AL_CH, M) - 1;
    double cp[AL_MAX], dd[3][AL_MAX];
    long long cur = j0;
    while (cur <= j1) {
        // segment extent around `cur`, limited to the halo
        long long a = cur; bool left_true = false;
        for (;;) {
            if (a == 0) { left_true = true; break; }
            const double st = gt[a] - gt[a - 1];
            if (st > gap) { left_true = true; break; }
            if (!(st > 1e-9)) *bad_steps = 1;              // not strictly increasing by > 1e-9: the reference drops the segment (:356-359)
            if (cur - a >= AL_H + (cur - j0)) break;
            --a;
        }
        long long b = cur; bool right_true = false;
        for (;;) {
            if (b == M - 1) { right_true = true; break; }
            const double st = gt[b + 1] - gt[b];
            if (st > gap) { right_true = true; break; }
            if (!(st > 1e-9)) *bad_steps = 1;
            if (b - j1 >= AL_H) break;
            ++b;
        }
        const long long w0 = max(a, j0), w1 = min(b, j1);   // knots of this chunk in the segment
        const int m = (int)(b - a + 1);
        if (m < 4 && left_true && right_true) {
            for (long long j = w0; j <= w1; ++j) { mom[3 * j] = 0.0; mom[3 * j + 1] = 0.0; mom[3 * j + 2] = 0.0; }
        } else {
            auto h = [&](long long j) { return gt[j + 1] - gt[j]; };
            // Thomas forward sweep over the interior unknowns a+1 .. b-1 (local index j - a); knot times, spacings and
            // slopes are carried from step to step: one new knot (t, x, y, z) is loaded per step
            double pcp = 0.0, pd[3] = {0.0, 0.0, 0.0};
            double tj = gt[a + 1], hp = tj - gt[a];                         // t_j, h(j-1)
            double yj[3] = {gy[3 * (a + 1)], gy[3 * (a + 1) + 1], gy[3 * (a + 1) + 2]};
            double sp[3] = {(yj[0] - gy[3 * a]) / hp, (yj[1] - gy[3 * a + 1]) / hp, (yj[2] - gy[3 * a + 2]) / hp};      // slope of interval j-1
            const double e_h0 = hp, e_h1 = gt[a + 2] - gt[a + 1];           // first two spacings (not-a-knot row at a true left end)
            const double e_ha = gt[b - 1] - gt[b - 2], e_hb = gt[b] - gt[b - 1];
            for (long long j = a + 1; j <= b - 1; ++j) {
                const double tn = gt[j + 1], hc = tn - tj;                 // h(j)
                double lo = hp, di = 2.0 * (hp + hc), up = hc;
                if (j == a + 1) {
                    if (left_true) { di += e_h0 * (e_h0 + e_h1) / e_h1; up -= e_h0 * e_h0 / e_h1; }
                    lo = 0.0;
                }
                if (j == b - 1) {
                    if (right_true) { di += e_hb * (e_ha + e_hb) / e_ha; lo -= e_hb * e_hb / e_ha; }
                    up = 0.0;
                }
                const double den = di - lo * pcp;
                const double c = up / den;
                const int L = (int)(j - a);
                cp[L] = c;
#pragma unroll
                for (int ax = 0; ax < 3; ++ax) {
                    const double yn = gy[3 * (j + 1) + ax];
                    const double sn = (yn - yj[ax]) / hc;
                    const double d = (6.0 * (sn - sp[ax]) - lo * pd[ax]) / den;
                    dd[ax][L] = d; pd[ax] = d;
                    yj[ax] = yn; sp[ax] = sn;
                }
                pcp = c; tj = tn; hp = hc;
            }
            for (long long j = b - 2; j >= a + 1; --j) {
                const int L = (int)(j - a);
#pragma unroll
                for (int ax = 0; ax < 3; ++ax) dd[ax][L] -= cp[L] * dd[ax][L + 1];
            }
#pragma unroll
            for (int ax = 0; ax < 3; ++ax) {
                if (left_true) { const double h0 = h(a), h1 = h(a + 1); dd[ax][0] = ((h0 + h1) * dd[ax][1] - h0 * dd[ax][2]) / h1; }
                else dd[ax][0] = 0.0;
                const int Lb = (int)(b - a);
                if (right_true) { const double ha = h(b - 2), hb = h(b - 1); dd[ax][Lb] = ((ha + hb) * dd[ax][Lb - 1] - hb * dd[ax][Lb - 2]) / ha; }
                else dd[ax][Lb] = 0.0;
            }
            for (long long j = w0; j <= w1; ++j) {
                const int L = (int)(j - a);
                mom[3 * j] = dd[0][L]; mom[3 * j + 1] = dd[1][L]; mom[3 * j + 2] = dd[2][L];
            }
        }
        cur = w1 + 1;
    }
}

__global__ void __launch_bounds__(256) assoc_long_eval_kernel(const double* __restrict__ gt, const double* __restrict__ gy, const double* __restrict__ mom,
                                                              long long M, const double* __restrict__ st, long long N, double gap,
                                                              double* __restrict__ out, unsigned char* __restrict__ val) {
    // The stamps of a block are usually close together (SLAM stamps are sorted): two threads bracket the block's smallest
    // and largest stamp in the whole knot array (27 dependent loads for 1e8 knots), every other thread searches inside that
    // bracket only (a few hundred knots, cached).
    __shared__ double s_lo[8], s_hi[8];
    __shared__ long long b_l, b_r;
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const double t = i < N ? st[i] : nan("");
    {
        double mn = (t == t) ? t : INFINITY, mx = (t == t) ? t : -INFINITY;
        for (int o = 16; o > 0; o >>= 1) { mn = fmin(mn, __shfl_xor_sync(GSF_FULL_MASK, mn, o)); mx = fmax(mx, __shfl_xor_sync(GSF_FULL_MASK, mx, o)); }
        if ((threadIdx.x & 31) == 0) { s_lo[threadIdx.x >> 5] = mn; s_hi[threadIdx.x >> 5] = mx; }
        __syncthreads();
        if (threadIdx.x < 2 && M >= 2) {
            double q = threadIdx.x == 0 ? INFINITY : -INFINITY;
            for (int w = 0; w < 8; ++w) q = threadIdx.x == 0 ? fmin(q, s_lo[w]) : fmax(q, s_hi[w]);
            long long l = 0, r = M - 1;                       // largest j with gt[j] <= q (0 if none), then widened by one
            while (r - l > 1) { const long long mid = (l + r) >> 1; if (gt[mid] <= q) l = mid; else r = mid; }
            if (threadIdx.x == 0) b_l = l; else b_r = r;
        }
        __syncthreads();
    }
    if (i >= N) return;
    double v0 = nan(""), v1 = v0, v2 = v0;
    if (M >= 2 && t >= gt[0] && t <= gt[M - 1]) {
        long long l = b_l, r = b_r;                           // largest j with gt[j] <= t: gt[b_l] <= t (or b_l = 0), gt[b_r] >= t (or b_r = M - 1)
        while (r - l > 1) { const long long mid = (l + r) >> 1; if (gt[mid] <= t) l = mid; else r = mid; }
        long long j = (gt[r] <= t) ? r : l;
        // interval [j, j+1] unless t sits exactly on the last knot of a segment: then [j-1, j]
        bool ok = true;
        if (j == M - 1 || gt[j + 1] - gt[j] > gap) {
            if (t == gt[j] && j > 0 && !(gt[j] - gt[j - 1] > gap)) --j; else ok = false;        // inside a gap (or an isolated knot): no segment
        }
        if (ok) {
            // knots of the segment around the interval, up to 4 (decides cubic / linear, :362)
            int cnt = 2;
            if (j >= 1 && !(gt[j] - gt[j - 1] > gap)) { ++cnt; if (j >= 2 && !(gt[j - 1] - gt[j - 2] > gap)) ++cnt; }
            if (j + 2 <= M - 1 && !(gt[j + 2] - gt[j + 1] > gap)) { ++cnt; if (j + 3 <= M - 1 && !(gt[j + 3] - gt[j + 2] > gap)) ++cnt; }
            const double hh = gt[j + 1] - gt[j];
            const double wa = (gt[j + 1] - t) / hh, wb = (t - gt[j]) / hh;
            double v[3];
#pragma unroll
            for (int ax = 0; ax < 3; ++ax) {
                const double y0 = gy[3 * j + ax], y1 = gy[3 * (j + 1) + ax];
                if (cnt >= 4) {
                    const double m0 = mom[3 * j + ax], m1 = mom[3 * (j + 1) + ax];
                    v[ax] = wa * y0 + wb * y1 + ((wa * wa * wa - wa) * m0 + (wb * wb * wb - wb) * m1) * (hh * hh) / 6.0;
                } else v[ax] = (y1 - y0) / hh * (t - gt[j]) + y0;
            }
            v0 = v[0]; v1 = v[1]; v2 = v[2];
        }
    }
    out[3 * i] = v0; out[3 * i + 1] = v1; out[3 * i + 2] = v2;
    val[i] = !row_has_nan(v0, v1, v2);
}

// work: 3 M doubles (moments) + 1 int
cudaError_t launch_associate_long(const double* gps_t, const double* gps_xyz, long long M, const double* slam_t, long long N, double gap,
                                  double* work, double* aligned, unsigned char* valid, int* status, cudaStream_t stream) {
    int* bad = reinterpret_cast<int*>(work + 3 * M);
    cudaError_t e = cudaMemsetAsync(bad, 0, sizeof(int), stream);
    if (e != cudaSuccess) return e;
    if (M >= 2) {
        const long long chunks = (M + AL_CH - 1) / AL_CH;
        assoc_long_moments_kernel<<<(unsigned)((chunks + 127) / 128), 128, 0, stream>>>(gps_t, gps_xyz, M, gap, work, bad);
    }
    if (N > 0) assoc_long_eval_kernel<<<(unsigned)((N + 255) / 256), 256, 0, stream>>>(gps_t, gps_xyz, work, M, slam_t, N, gap, aligned, valid);
    e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    if (status) e = cudaMemcpyAsync(status, bad, sizeof(int), cudaMemcpyDeviceToDevice, stream);
    return e;
}

}  // namespace gsf
