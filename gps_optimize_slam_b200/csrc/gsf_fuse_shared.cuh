// Device helpers shared by the two fused-path kernels (gsf_fused.cu: general kernel, gsf_fast.cu:
// warp-specialised fast kernel): fast reciprocal, Moebius / affine scan operators, the TMA trajectory
// load and the streaming quaternion pass.
#pragma once
#include "gsf_common.cuh"
#include "gsf_ekf_strict.cuh"
#include "gsf_ptx.cuh"
#include "gsf_internal.cuh"

namespace gsf {

// 1/x for finite positive x: hardware seed (rel. error 2^-23) + two Newton steps -> <= 1 ulp.
// 5 instructions instead of the ~10 of the IEEE division sequence; parity budget is 1e-9.
__device__ __forceinline__ double fast_rcp(double x) {
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    double e = fma(-x, r, 1.0);
    r = fma(r, e, r);
    e = fma(-x, r, 1.0);
    return fma(r, e, r);
}

// 1/sqrt(x) for positive normal x: hardware seed + one third-order step (the arithmetic of CUDA's rsqrt() without
// its special-case branch, so that two chains interleave in one loop body).  x = 0 gives inf (-> NaN downstream).
__device__ __forceinline__ double fast_rsqrt(double x) {
    double r;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    const double e = fma(-x, r * r, 1.0);
    return fma(fma(e, 0.375, 0.5), r * e, r);
}

constexpr int FLAG_VALID = 1;
constexpr int FLAG_SELECTED = 2;
constexpr int FLAG_RECOVERY = 4;
constexpr int FLAG_NO_RTS = 8;
constexpr int FLAG_SHARP_STEP = 16;     // outage step (pose i-1 -> i, both without GNSS) whose yaw rate exceeds the sharp-turn threshold

// ----------------------------------------------------------------------------- scan operators
// 2x2 Moebius matrices for the three position axes, row-major [a b; c d] per axis.  Entries
// are non-negative (no cancellation); rescaled by a power of two so products never under- or
// overflow however long the trajectory is.
struct Moeb3 { double m[12]; };
struct Aff3 { double a[3], b[3]; };                     // x -> a x + b per axis

__device__ __forceinline__ void moeb_rescale(double* m) {
    // entries are >= 0, so the largest one has the largest high word: three integer maxima instead of three fmax
    const int hi = max(max(__double2hiint(m[0]), __double2hiint(m[1])), max(__double2hiint(m[2]), __double2hiint(m[3])));
    const int e = ((hi >> 20) & 0x7ff) - 1023;
    const double sc = __hiloint2double((1023 - e) << 20, 0);       // 2^-e, exact
    m[0] *= sc; m[1] *= sc; m[2] *= sc; m[3] *= sc;
}
__device__ __forceinline__ void moeb_identity(Moeb3& x) {
#pragma unroll
    for (int a = 0; a < 3; ++a) { x.m[4 * a] = 1.0; x.m[4 * a + 1] = 0.0; x.m[4 * a + 2] = 0.0; x.m[4 * a + 3] = 1.0; }
}
// r = later o earlier.  NAX == 2: only axes 0 and 2 are carried (axis 1 is a copy of axis 0).
template <int NAX>
__device__ __forceinline__ Moeb3 moeb_compose(const Moeb3& e, const Moeb3& l) {
    Moeb3 r;
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        if (NAX == 2 && a == 1) continue;
        const double* E = e.m + 4 * a; const double* L = l.m + 4 * a; double* R = r.m + 4 * a;
        R[0] = L[0] * E[0] + L[1] * E[2]; R[1] = L[0] * E[1] + L[1] * E[3];
        R[2] = L[2] * E[0] + L[3] * E[2]; R[3] = L[2] * E[1] + L[3] * E[3];
        moeb_rescale(R);
    }
    return r;
}
__device__ __forceinline__ Aff3 aff_compose(const Aff3& e, const Aff3& l) {
    Aff3 r;
#pragma unroll
    for (int a = 0; a < 3; ++a) { r.a[a] = l.a[a] * e.a[a]; r.b[a] = l.a[a] * e.b[a] + l.b[a]; }
    return r;
}
// Inclusive warp scans (fixed shuffle pattern => deterministic).
template <int NAX>
__device__ __forceinline__ void moeb_warp_scan(Moeb3& x, int lane) {
#pragma unroll 1
    for (int o = 1; o < 32; o <<= 1) {
        Moeb3 y;
#pragma unroll
        for (int k = 0; k < 12; ++k) {
            if (NAX == 2 && k >= 4 && k < 8) continue;
            y.m[k] = __shfl_up_sync(GSF_FULL_MASK, x.m[k], o);
        }
        if (lane >= o) x = moeb_compose<NAX>(y, x);
    }
}
// Chunk composite of the covariance maps of steps [s0, c1) and its warp scan.  Returns the
// warp-exclusive prefix; the warp total goes to `wtot` (lane 31).
template <int NAX>
__device__ __forceinline__ Moeb3 moebius_chunk_scan(const double* tsS, const unsigned char* flg, const FuseParams& prm,
                                                    int s0, int c1, int lane, double* wtot) {
    Moeb3 loc;
    moeb_identity(loc);
    int since = 0;
    for (int i = s0; i < c1; ++i) {
        const double dt = fmax(1e-6, tsS[i] - tsS[i - 1]);
        const bool v = flg[i] & FLAG_VALID;
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            if (NAX == 2 && a == 1) continue;
            double* m = loc.m + 4 * a;
            const double qa = prm.q[a] * dt, ra = prm.r[a];
            const double ta = m[0] + qa * m[2], tb = m[1] + qa * m[3];        // [1 q; 0 1] * m
            if (v) { m[2] = ta + ra * m[2]; m[3] = tb + ra * m[3]; m[0] = ra * ta; m[1] = ra * tb; }
            else { m[0] = ta; m[1] = tb; }
        }
        if (++since == 16) {
            since = 0;
            moeb_rescale(loc.m); if (NAX == 3) moeb_rescale(loc.m + 4); moeb_rescale(loc.m + 8);
        }
    }
    moeb_rescale(loc.m); if (NAX == 3) moeb_rescale(loc.m + 4); moeb_rescale(loc.m + 8);
    moeb_warp_scan<NAX>(loc, lane);
    if (NAX == 2) { loc.m[4] = loc.m[0]; loc.m[5] = loc.m[1]; loc.m[6] = loc.m[2]; loc.m[7] = loc.m[3]; }
    if (lane == 31) {
#pragma unroll
        for (int k = 0; k < 12; ++k) wtot[k] = loc.m[k];
    }
    Moeb3 mex;
#pragma unroll
    for (int k = 0; k < 12; ++k) mex.m[k] = __shfl_up_sync(GSF_FULL_MASK, loc.m[k], 1);
    if (lane == 0) moeb_identity(mex);
    return mex;
}
__device__ __forceinline__ void aff_warp_scan(Aff3& x, int lane) {
#pragma unroll 1
    for (int o = 1; o < 32; o <<= 1) {
        Aff3 y;
#pragma unroll
        for (int k = 0; k < 3; ++k) { y.a[k] = __shfl_up_sync(GSF_FULL_MASK, x.a[k], o); y.b[k] = __shfl_up_sync(GSF_FULL_MASK, x.b[k], o); }
        if (lane >= o) x = aff_compose(y, x);
    }
}

// Rare / once-per-trajectory code kept out of line so that its register needs do not spill
// the per-pose loops.
static __device__ __noinline__ int umeyama_finish_ool(int n, const double* mu_s, const double* mu_d, const double* H, double ss,
                                               double* R, double* t, double* s) {
    return umeyama_finish(n, mu_s, mu_d, H, ss, R, t, *s);
}
static __device__ __noinline__ bool sharp_turn_ool(const double* ts, const double* quat, long s, long e, double thresh) {
    return sharp_turn_in_range(ts, quat, s, e, thresh);
}


// one thread: TMA bulk copies of trajectory b (even element count from an even start) and an L2
// prefetch of its quaternions.
__device__ __forceinline__ void issue_trajectory_load(const FuseArgs& A, int b, double* ts_s, double* pos_s, double* z_s, uint64_t* mbar) {
    const long long e0 = A.offsets[b];
    const int n = (int)(A.offsets[b + 1] - e0);
    if (n <= 0 || n > A.cap) return;
    if (A.only_deferred && A.status[b] != ST_DEFERRED) return;
    const int lead = (int)(e0 & 1);
    const int even = (n + lead) & ~1;
    if (even > 0) {
        mbar_expect_tx(mbar, (uint32_t)even * 56u);
        bulk_g2s(ts_s, A.ts + (e0 - lead), (uint32_t)even * 8u, mbar);
        bulk_g2s(pos_s, A.pos + 3 * (e0 - lead), (uint32_t)even * 24u, mbar);
        bulk_g2s(z_s, A.z + 3 * (e0 - lead), (uint32_t)even * 24u, mbar);
    }
    // its quaternions are streamed later (global -> global): warm L2 now so that pass hits L2
    const long long qn = ((long long)n * 32) & ~15ll;
    if (qn > 0) bulk_prefetch_l2(A.quat + 4 * e0, (uint32_t)qn);
}

// Streaming quaternion pass: q_state[i] = C (x) q_hat[i] for poses first, first+stride, ... < n,
// 4 poses in flight per thread.  Returns 1 if a zero-norm quaternion was met.
__device__ __forceinline__ int quat_rounds(const double* __restrict__ quat_in, double* __restrict__ quat_out, const Quat& C,
                                           int first, int stride, int n) {
    const double2* __restrict__ qin = reinterpret_cast<const double2*>(quat_in);
    double2* __restrict__ qout = reinterpret_cast<double2*>(quat_out);
    int bad = 0;
    for (int i0 = first; i0 < n; i0 += 4 * stride) {
        double2 lo[4], hi[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int i = i0 + u * stride;
            if (i < n) { lo[u] = __ldg(qin + 2 * i); hi[u] = __ldg(qin + 2 * i + 1); }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int i = i0 + u * stride;
            if (i < n) {
                const Quat qi{lo[u].x, lo[u].y, hi[u].x, hi[u].y};
                const double n2 = qnorm2(qi);
                if (n2 == 0.0) bad = 1;                     // scipy raises here (:466); output row becomes NaN
                const Quat r = qscale(qmul(C, qi), rsqrt(n2));
                qout[2 * i] = make_double2(r.x, r.y);
                qout[2 * i + 1] = make_double2(r.z, r.w);
            }
        }
    }
    return bad;
}


}  // namespace gsf
