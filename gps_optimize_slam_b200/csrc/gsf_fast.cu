// gsf_fuse_batched, fast kernel: the same fused Sim3 -> EKF path as gsf_fused.cu for the common case
// (every pose has a GNSS measurement, no GNSS time gap, trajectory inside the Sim3 window), organised as
// a warp-specialised pipeline so that nothing serial sits on the critical path of the pose loops.
//
// Reference path replaced (file:line in /root/reference/EKFGPSSLAM.py): Sim3 point selection :972-998
// (here: "all points", verified on the fly), compute_sim3_transform :428-459, transform_trajectory row 0
// :461-467, apply_ekf_correction :831-935 with ExtendedKalmanFilter :679-772.  Trajectories that need
// the general machinery (invalid rows -> outages/RTS :875-928, time gaps, Sim3 window, too few points,
// zero quaternion) are detected by the look-ahead warps, marked ST_DEFERRED and processed by the general
// kernel (gsf_fused.cu) launched behind this one on the same stream.
//
// One block = CT compute threads + two look-ahead warps, persistent; trajectories are handed out by a global
// counter through a four-entry queue in shared memory (compute thread 0 requests entry j + 3 while trajectory j
// runs), so that every block ends within one trajectory of the others whatever its pace:
//   * warp A (sums)  streams positions + measurements of trajectory j+1/j+2 straight from HBM with
//     128-bit loads (the later TMA load of the same bytes hits L2), accumulates the 16 pivot-shifted
//     Umeyama sums in a fixed order (lane-strided pose pairs, transposing butterfly) -> bit-reproducible;
//   * warp B (scan + SVD) reads the timestamps, checks gap/window, composes the covariance recursion
//     p -> R(p+q)/(p+q+R) per compute-thread chunk as 2x2 Moebius matrices [1 qa; g g*qa+1], g = 1/R (branch-free
//     interleaved chains, one warp scan over 32 lanes x CT/32 chunks) and publishes the covariance every compute
//     thread starts from; then finishes Umeyama (one-sided Jacobi SVD) and publishes R, t, s,
//     C = q_state0 (x) conj(q_hat0), M(C);
//   * compute warps (trajectory j, resident in shared memory via three TMA bulk copies): pass B (telescoped
//     odometry, residual check, per-step gains from the projectively carried covariance, affine state maps),
//     affine scan, state recursion (pass C), bulk store, quaternion pass (TMA in two parts over the dead
//     timestamp + position buffers), next TMA load.  No Moebius scan, no reduction, no SVD wait.
//   Hand-offs are hardware named barriers (aux_ready / aux_free / sums_ready, two slots: bar.arrive on the
//   signalling side, bar.sync on the waiting side); mbarriers only for the TMA completions.
// The kernel is instruction-supply bound (12 resident warps per SM in three roles, ~85 KB of hot code against a
// 32 KB L1.5 instruction cache): every per-pose loop is rolled, the roles are inlined into the kernel (parameters
// from the constant bank), debug stamps are compiled out (GSF_DEBUG_STAMPS).  See DESIGN.md section 3.
#include <cstdlib>
#include <atomic>
#include "gsf_fuse_shared.cuh"

namespace gsf {

#ifndef GSF_PASSC_UNROLL
#define GSF_PASSC_UNROLL 1
#endif
#ifndef GSF_A_EARLY
#define GSF_A_EARLY 0               // 0: warp A starts trajectory j+2 when trajectory j releases its slot; 1 / 2: when trajectory j has finished pass B / pass C
#endif
#ifndef GSF_SUMS_SPLIT
#define GSF_SUMS_SPLIT 0            // 1: the scan warp sums the last quarter of the pose pairs (see sums_share_b; measured neutral)
#endif
#ifndef GSF_QUAT2
#define GSF_QUAT2 0                 // quaternion pass: poses per thread and iteration - 1
#endif
#ifndef GSF_PASS_UNROLL
#define GSF_PASS_UNROLL 1
#endif
constexpr int PASSC_UNROLL = GSF_PASSC_UNROLL;
constexpr int PASS_UNROLL = GSF_PASS_UNROLL;      // per-pose loops of the compute warps stay rolled: with the roles inlined, one step per
                                    // iteration measured 3 % faster than two and 8 % faster than four (instruction-cache footprint)


#ifdef GSF_DEBUG_STAMPS                             // tools/phase_timing_fast.py: build with GSF_NVCC_EXTRA=-DGSF_DEBUG_STAMPS
#define GSF_FSTAMP(k) do { if (pclk) { const long long t_ = clock64(); pclk[k] += t_ - tlast; tlast = t_; } } while (0)   // pclk: per-role debug pointer; cycles since the role's previous stamp, accumulated over the block's trajectories
#define GSF_FSTAMP_DECL long long tlast = clock64(); (void)tlast
#else
#define GSF_FSTAMP(k) do { } while (0)
#define GSF_FSTAMP_DECL do { } while (0)
#endif

// ----------------------------------------------------------------------------- Moebius maps, NAX axes
template <int NAX> struct MoebN { double m[4 * NAX]; };      // per axis row-major [a b; c d]

template <int NAX>
__device__ __forceinline__ void moebn_identity(MoebN<NAX>& x) {
#pragma unroll
    for (int a = 0; a < NAX; ++a) { x.m[4 * a] = 1.0; x.m[4 * a + 1] = 0.0; x.m[4 * a + 2] = 0.0; x.m[4 * a + 3] = 1.0; }
}
template <int NAX>
__device__ __forceinline__ void moebn_rescale(MoebN<NAX>& x) {
#pragma unroll
    for (int a = 0; a < NAX; ++a) moeb_rescale(x.m + 4 * a);
}
// r = later o earlier
template <int NAX>
__device__ __forceinline__ MoebN<NAX> moebn_compose(const MoebN<NAX>& e, const MoebN<NAX>& l) {
    MoebN<NAX> r;
#pragma unroll
    for (int a = 0; a < NAX; ++a) {
        const double* E = e.m + 4 * a; const double* L = l.m + 4 * a; double* R = r.m + 4 * a;
        R[0] = L[0] * E[0] + L[1] * E[2]; R[1] = L[0] * E[1] + L[1] * E[3];
        R[2] = L[2] * E[0] + L[3] * E[2]; R[3] = L[2] * E[1] + L[3] * E[3];
        moeb_rescale(R);
    }
    return r;
}
// r = later o earlier, no rescale (callers rescale every other scan stage: entries are >= 0 and every operand was
// normalised to a maximum in [1, 2) at most two products ago, so neither overflow nor harmful underflow can occur)
template <int NAX>
__device__ __forceinline__ MoebN<NAX> moebn_compose_raw(const MoebN<NAX>& e, const MoebN<NAX>& l) {
    MoebN<NAX> r;
#pragma unroll
    for (int a = 0; a < NAX; ++a) {
        const double* E = e.m + 4 * a; const double* L = l.m + 4 * a; double* R = r.m + 4 * a;
        R[0] = L[0] * E[0] + L[1] * E[2]; R[1] = L[0] * E[1] + L[1] * E[3];
        R[2] = L[2] * E[0] + L[3] * E[2]; R[3] = L[2] * E[1] + L[3] * E[3];
    }
    return r;
}
// one predict + update step with a valid measurement, applied on the left.  The step matrix [r r*q; 1 q+r] is
// used divided by r (a Moebius map is projective): [1 qa; g g*qa+1] with g = 1/r -- four FMAs per axis, and
// (qa, g) = (0, 0) is the identity, so steps outside the trajectory are masked by selecting the two scalars
// instead of branching (the chains of one lane then interleave in one straight-line loop body).
template <int NAX>
__device__ __forceinline__ void moebn_step(MoebN<NAX>& x, const double* qv, const double* gv, double dt, bool valid) {
#pragma unroll
    for (int a = 0; a < NAX; ++a) {
        double* m = x.m + 4 * a;
        const double qa = qv[a] * dt, ga = valid ? gv[a] : 0.0;
        m[0] = fma(qa, m[2], m[0]); m[1] = fma(qa, m[3], m[1]);
        m[2] = fma(ga, m[0], m[2]); m[3] = fma(ga, m[1], m[3]);
    }
}
template <int NAX>
__device__ __forceinline__ double moebn_apply(const MoebN<NAX>& x, int a, double p) {
    const double* m = x.m + 4 * a;
    return (m[0] * p + m[1]) * rcp_(m[2] * p + m[3]);      // positive normal denominator (entries >= 0, rescaled)
}

// ----------------------------------------------------------------------------- shared-memory map (doubles after the staging buffers)
constexpr int FS_BC = 0;            // 2 slots x 48: M(C) 0-8, C 9-12, x0 13-15, t 16-18, s 19, R 20-28, verdict 30, status 31
constexpr int FS_SUMS = 96;         // 2 slots x 24: 16 sums, pivots 16-21
constexpr int FS_AFF = 144;         // up to 8 warps x 6 affine warp totals (48)
constexpr int FS_PRM = 192;         // FuseParams as 23 doubles (24): compute warps' copy
constexpr int FS_INT = 216;         // 4 ints: 0 residual violators
constexpr int FS_MBAR = 218;        // 8 mbarriers: full, sums_ready[2], aux_ready[2], aux_free[2], ts_b
constexpr int FS_PRMB = 226;        // FuseParams, warp B's copy (24)
constexpr int FS_RING = 250;        // 4 trajectory references (16 bytes each): the block's work queue, filled by compute thread 0
constexpr int FS_PST = 258;         // 2 slots x CT x 3 start covariances; then warp B's timestamp buffer (cap2 doubles)
constexpr int MB_FULL = 0, MB_QUAT = 1, MB_QUAT0 = 2, MB_TSB = 7;
// Role hand-offs use hardware named barriers (bar.sync on the waiting side, bar.arrive on the signalling side): a
// parked warp costs no issue slots (warps polling an mbarrier slowed the serial SVD of the warp they were waiting
// for).  Barrier ids (two slots each): 1 compute-internal, 2-3 aux_ready, 4-5 / 6-7 aux_free for the sums / scan
// warp, 8-9 sums_ready.  mbarriers remain for the TMA completions only (MB_FULL, MB_TSB; polled by one lane with
// nanosleep back-off).
constexpr int NB_AUXRDY = 2, NB_FREE_A = 4, NB_FREE_B = 6, NB_SUMS = 8, NB_MID = 10;

__host__ __device__ constexpr size_t fast_smem_bytes(int cap, int ct) {
    return (size_t)((cap + 3) & ~1) * 64 + (size_t)(FS_PST + 6 * ct + 16) * 8;      // ... + 16: the scan warp's partial sums
}

// 16 accumulators x 32 lanes -> lane L (bit 0 clear) ends with the warp total of value idx(L):
// fixed exchange pattern (halve the value set at each of the first four stages), 16 shuffles of doubles.
__device__ __forceinline__ double butterfly16(double* v, int lane) {
#pragma unroll
    for (int st = 0; st < 4; ++st) {
        const int o = 16 >> st, half = 8 >> st;
        const bool up = lane & o;
#pragma unroll
        for (int k = 0; k < half; ++k) {
            const double send = up ? v[k] : v[k + half];
            const double keep = up ? v[k + half] : v[k];
            v[k] = keep + __shfl_xor_sync(GSF_FULL_MASK, send, o);
        }
    }
    return v[0] + __shfl_xor_sync(GSF_FULL_MASK, v[0], 1);
}
__device__ __forceinline__ int butterfly16_index(int lane) { return ((lane >> 4) & 1) * 8 + ((lane >> 3) & 1) * 4 + ((lane >> 2) & 1) * 2 + ((lane >> 1) & 1); }

__device__ __forceinline__ void umeyama_accumulate(double* v, double p0, double p1, double p2, double z0, double z1, double z2,
                                                   double ps0, double ps1, double ps2, double pz0, double pz1, double pz2) {
    const double a0 = p0 - ps0, a1 = p1 - ps1, a2 = p2 - ps2;
    const double b0 = z0 - pz0, b1 = z1 - pz1, b2 = z2 - pz2;
    v[0] += a0; v[1] += a1; v[2] += a2; v[3] += b0; v[4] += b1; v[5] += b2;
    v[6] += a0 * b0; v[7] += a0 * b1; v[8] += a0 * b2;
    v[9] += a1 * b0; v[10] += a1 * b1; v[11] += a1 * b2;
    v[12] += a2 * b0; v[13] += a2 * b1; v[14] += a2 * b2;
    v[15] += a0 * a0 + a1 * a1 + a2 * a2;
}

// Warp B, part 1: covariance the compute thread t starts from, for every t, and the gap / window check.
template <int NAX, int CT, int LCH>
__device__ __forceinline__ int cov_start_scan(const double* __restrict__ gts, int n, const FuseParams* __restrict__ gp, int lane,
                                              double* __restrict__ pst, long long* clk) {
    constexpr int CPL = CT / 32;                            // compute-thread chunks per lane
    double qv[NAX], gv[NAX], p0v[NAX];
    int viol = 0;
#pragma unroll
    for (int a = 0; a < NAX; ++a) {
        const int ax = (NAX == 2 && a == 1) ? 2 : a;
        qv[a] = gp->q[ax]; p0v[a] = gp->p0[ax];
        const double ra = gp->r[ax];
        if (!(ra >= 1e-300 && ra <= 1e300)) viol = 1;        // exact (zero-variance) or non-finite measurement noise: general kernel
        gv[a] = 1.0 / ra;
    }
    const double gap = gp->gap_threshold, t_lim = gts[0] + gp->max_duration;
    MoebN<NAX> incl[CPL - 1 > 0 ? CPL - 1 : 1];
    MoebN<NAX> cur;
    moebn_identity(cur);
    // The lane's CPL chunks are independent 2x2 product chains: advance them together, one step each per
    // iteration (CPL-way instruction-level parallelism, branch-free; the loop stays rolled), then compose them in order.
    MoebN<NAX> ch[CPL];
    double tpk[CPL];
#pragma unroll
    for (int k = 0; k < CPL; ++k) {
        moebn_identity(ch[k]);
        const int c0 = min((lane * CPL + k) * LCH, n);
        tpk[k] = c0 >= 1 && c0 < n ? gts[c0 - 1] : (n > 0 ? gts[0] : 0.0);
    }
#pragma unroll 1
    for (int m = 0; m < LCH; ++m) {
#pragma unroll
        for (int k = 0; k < CPL; ++k) {
            const int i = (lane * CPL + k) * LCH + m;
            const bool valid = i >= 1 && i < n;
            const double ti = gts[min(i, n - 1)];
            const double raw = ti - tpk[k];
            if (valid && (!(raw > 1e-6) || raw > gap || ti > t_lim)) viol = 1;     // dt <= 1e-6 s (the reference clamps, :863) or NaN: general kernel
            moebn_step(ch[k], qv, gv, valid ? raw : 0.0, valid);
            tpk[k] = ti;
        }
    }
#pragma unroll
    for (int k = 0; k < CPL; ++k) {
        double fin = 0.0;
#pragma unroll
        for (int q = 0; q < 4 * NAX; ++q) fin += ch[k].m[q];
        if (!(fin <= 1e200)) viol = 1;                       // q/r ratio beyond the range of a chunk product (here and in pass B2): general kernel
        moebn_rescale(ch[k]);
        cur = k == 0 ? ch[0] : moebn_compose(cur, ch[k]);   // inclusive prefix inside the lane
        if (k < CPL - 1) incl[k] = cur;
    }
    if (clk) clk[29] = clock64();
    // inclusive warp scan of the lane totals, then the exclusive prefix of this lane
    MoebN<NAX> tot = cur;
#pragma unroll 1
    for (int o = 1; o < 32; o <<= 1) {
        MoebN<NAX> y;
#pragma unroll
        for (int k = 0; k < 4 * NAX; ++k) y.m[k] = __shfl_up_sync(GSF_FULL_MASK, tot.m[k], o);
        if (lane >= o) tot = moebn_compose_raw(y, tot);
        if (o == 2 || o == 8 || o == 16) moebn_rescale(tot);           // every other stage (and the last)
    }
    MoebN<NAX> ex;
#pragma unroll
    for (int k = 0; k < 4 * NAX; ++k) ex.m[k] = __shfl_up_sync(GSF_FULL_MASK, tot.m[k], 1);
    if (lane == 0) moebn_identity(ex);
    if (clk) clk[30] = clock64();
    double pl[NAX];
#pragma unroll
    for (int a = 0; a < NAX; ++a) pl[a] = moebn_apply(ex, a, p0v[a]);
#pragma unroll
    for (int k = 0; k < CPL; ++k) {
        const int t = lane * CPL + k;
        double pk[NAX];
#pragma unroll
        for (int a = 0; a < NAX; ++a) pk[a] = k == 0 ? pl[a] : moebn_apply(incl[k > 0 ? k - 1 : 0], a, pl[a]);
        if (NAX == 2) { pst[3 * t] = pk[0]; pst[3 * t + 1] = pk[0]; pst[3 * t + 2] = pk[1]; }
        else { pst[3 * t] = pk[0]; pst[3 * t + 1] = pk[1]; pst[3 * t + 2] = pk[NAX - 1]; }
    }
    return __any_sync(GSF_FULL_MASK, viol);
}

// one thread: TMA bulk copies of a trajectory (second and last touch of these bytes: evict_first)
__device__ __forceinline__ void issue_trajectory_load_hint(const FuseArgs& A, long long e0, int n, double* ts_s, double* pos_s, double* z_s, uint64_t* mbar) {
    if (n <= 0 || n > A.cap) return;
    const uint64_t pf = l2_policy_evict_first();
    const int lead = (int)(e0 & 1);
    const int even = (n + lead) & ~1;
    if (even > 0) {
        mbar_expect_tx(mbar, (uint32_t)even * 56u);
        bulk_g2s_hint(ts_s, A.ts + (e0 - lead), (uint32_t)even * 8u, mbar, pf);
        bulk_g2s_hint(pos_s, A.pos + 3 * (e0 - lead), (uint32_t)even * 24u, mbar, pf);
        bulk_g2s_hint(z_s, A.z + 3 * (e0 - lead), (uint32_t)even * 24u, mbar, pf);
    }
}
// Streaming quaternion pass q_state[i] = C (x) q_hat[i] (see quat_rounds) with evict_first loads and
// stores, U poses in flight per thread.
template <int U>
__device__ __forceinline__ int quat_rounds_hint(const double* __restrict__ quat_in, double* __restrict__ quat_out, const Quat& C,
                                                int first, int stride, int n) {
    const double2* __restrict__ qin = reinterpret_cast<const double2*>(quat_in);
    double2* __restrict__ qout = reinterpret_cast<double2*>(quat_out);
    const uint64_t pf = l2_policy_evict_first();
    int bad = 0;
#pragma unroll 1
    for (int i0 = first; i0 < n; i0 += U * stride) {
        double2 lo[U], hi[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int i = i0 + u * stride;
            if (i < n) { lo[u] = ldg2_hint(qin + 2 * i, pf); hi[u] = ldg2_hint(qin + 2 * i + 1, pf); }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int i = i0 + u * stride;
            if (i < n) {
                const Quat qi{lo[u].x, lo[u].y, hi[u].x, hi[u].y};
                const double n2 = qnorm2(qi);
                if (n2 == 0.0) bad = 1;                     // scipy raises here (:466); output row becomes NaN
                const Quat r = qscale(qmul(C, qi), rsqrt(n2));
                stg2_hint(qout + 2 * i, make_double2(r.x, r.y), pf);
                stg2_hint(qout + 2 * i + 1, make_double2(r.z, r.w), pf);
            }
        }
    }
    return bad;
}

// warp B, lane 0: bulk copy of the timestamps of trajectory b into warp B's private buffer
__device__ __forceinline__ void issue_ts_load(const FuseArgs& A, long long e0, int n, double* tsb, uint64_t* bar) {
    const int lead = (int)(e0 & 1), even = (n + lead) & ~1;
    mbar_expect_tx(bar, (uint32_t)even * 8u);
    bulk_g2s_hint(tsb, A.ts + (e0 - lead), (uint32_t)even * 8u, bar, l2_policy_evict_last());
}
// L2 prefetch of positions + measurements of a trajectory (issued by warp A right before it streams them)
__device__ __forceinline__ void prefetch_pos_z(const FuseArgs& A, long long e0, int n) {
    const long long lead = e0 & 1;
    const uint32_t bytes = (uint32_t)(((long long)n + lead) * 24) & ~15u;
    const uint64_t pl = l2_policy_evict_last();
    if (bytes) { bulk_prefetch_l2_hint(A.pos + 3 * (e0 - lead), bytes, pl); bulk_prefetch_l2_hint(A.z + 3 * (e0 - lead), bytes, pl); }
}

// Work distribution: trajectories are handed out by a global counter (one atomicAdd per trajectory, issued three
// trajectories ahead by compute thread 0), not by a static stride, so that all blocks finish within one trajectory
// of each other whatever their individual pace.  A reference with b >= B ends the block's sequence.
struct __align__(16) TrajRef { long long e0; int b; int n; };
__device__ __forceinline__ TrajRef traj_ref_from(const FuseArgs& A, int b) {      // b: value returned by the counter
    for (;;) {
        if (b >= A.B) return TrajRef{0, A.B, 0};
        const long long o0 = A.offsets[b], o1 = A.offsets[b + 1];
        const long long n = o1 - o0;
        if (n > 0 && n <= A.cap) return TrajRef{o0, b, (int)n};
        A.status[b] = n <= 0 ? ST_EMPTY : ST_TOO_LONG;     // not handled by this kernel: take another one
        b = atomicAdd(A.work_counter, 1);
    }
}

constexpr int fast_min_blocks(int ct) { return ct <= 32 ? 5 : (ct == 64 ? 3 : (ct <= 128 ? 3 : 2)); }

// Pass B of the compute warps, steps [s0, c1) of one thread, in one loop:
//  * telescoped odometry u_i = M(C)(p_i - p_{i-1}) and the residual of the Sim3 image y_i = s M(C) p_i + t against
//    the measurement (y advanced by s*u each step);
//  * per-step gains from the start covariance and the affine map x -> om x + (om u + k z) per axis; om overwrites
//    the position row, om u + k z the measurement row.  The covariance is carried projectively, P = a / b, with the
//    same step matrix as the scan warp ([1 qa; g g*qa+1], g = 1/r):  a' = a + qa b,  b' = b + g a',  k = g a' / b',
//    om = 1 - k (one FMA).  The recursion (two dependent operations per step) is separated from the reciprocal, which
//    pipelines across steps; k and om keep full relative accuracy.  P' = a'/b' = r pp / (pp + r) equals the
//    reference's Joseph form (:731) up to rounding, and the recursion is contractive, so the difference stays at the
//    1e-16 level.  XY: x and y share P0/Q/R (the shipped CONFIG), so the y gain is the x gain.
template <bool XY>
__device__ __forceinline__ int pass_b12(const double* __restrict__ tsS, double* __restrict__ posS, double* __restrict__ zS,
                                        const double* __restrict__ bc, const FuseParams& prm, const double* __restrict__ pst,
                                        int s0, int c1, double thr2, double pprev0, double pprev1, double pprev2, Aff3& aff) {
    double RC[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) RC[k] = bc[k];
    const double sc = bc[19];
    int nviol = 0;
    double y0 = 0.0, y1 = 0.0, y2 = 0.0;
    if (thr2 > 0.0 && s0 < c1) {
        mat_vec(RC, pprev0, pprev1, pprev2, y0, y1, y2);
        y0 = sc * y0 + bc[16]; y1 = sc * y1 + bc[17]; y2 = sc * y2 + bc[18];
    }
#pragma unroll
    for (int a = 0; a < 3; ++a) { aff.a[a] = 1.0; aff.b[a] = 0.0; }
    double ax = pst[0], ay = pst[1], az = pst[2], bx = 1.0, by = 1.0, bz = 1.0;
    const double qx = prm.q[0], qy = prm.q[1], qz = prm.q[2];
    const double gx = fast_rcp(prm.r[0]), gy = XY ? gx : fast_rcp(prm.r[1]), gz = fast_rcp(prm.r[2]);   // r: positive normal (scan warp defers otherwise)
#pragma unroll PASS_UNROLL
    for (int i = s0; i < c1; ++i) {
        const double p0 = posS[3 * i], p1 = posS[3 * i + 1], p2 = posS[3 * i + 2];
        const double z0 = zS[3 * i], z1 = zS[3 * i + 1], z2 = zS[3 * i + 2];
        // the reference's max(1e-6, dt) (:863) never binds here: the scan warp defers every trajectory with a
        // step of 1e-6 s or less (or NaN) to the general kernel
        const double dt = tsS[i] - tsS[i - 1];
        double u0, u1, u2;
        mat_vec(RC, p0 - pprev0, p1 - pprev1, p2 - pprev2, u0, u1, u2);
        pprev0 = p0; pprev1 = p1; pprev2 = p2;
        {   // unconditional (the count is dropped below when the check is off): the loop body stays one basic block, so
            // that consecutive steps can be interleaved by the instruction scheduler
            y0 = fma(sc, u0, y0); y1 = fma(sc, u1, y1); y2 = fma(sc, u2, y2);
            const double d0 = y0 - z0, d1 = y1 - z1, d2 = y2 - z2;
            nviol += (d0 * d0 + d1 * d1 + d2 * d2 < thr2) ? 0 : 1;
        }
        // in-place recursion (no loop-carried copies): a += qa b; ga = g a; b += ga; k = ga / b; om = 1 - k
        double kx, ky, kz, ox, oy, oz;
        {
            ax = fma(qx * dt, bx, ax);
            const double ga = gx * ax;
            bx += ga;
            const double rb = fast_rcp(bx);
            kx = ga * rb; ox = fma(-ga, rb, 1.0);
        }
        if (XY) { ky = kx; oy = ox; }
        else {
            ay = fma(qy * dt, by, ay);
            const double ga = gy * ay;
            by += ga;
            const double rb = fast_rcp(by);
            ky = ga * rb; oy = fma(-ga, rb, 1.0);
        }
        {
            az = fma(qz * dt, bz, az);
            const double ga = gz * az;
            bz += ga;
            const double rb = fast_rcp(bz);
            kz = ga * rb; oz = fma(-ga, rb, 1.0);
        }
        const double b0 = ox * u0 + kx * z0, b1 = oy * u1 + ky * z1, b2 = oz * u2 + kz * z2;
        posS[3 * i] = ox; posS[3 * i + 1] = oy; posS[3 * i + 2] = oz;
        zS[3 * i] = b0; zS[3 * i + 1] = b1; zS[3 * i + 2] = b2;
        aff.b[0] = ox * aff.b[0] + b0; aff.b[1] = oy * aff.b[1] + b1; aff.b[2] = oz * aff.b[2] + b2;
        aff.a[0] *= ox; aff.a[1] *= oy; aff.a[2] *= oz;
    }
    return thr2 > 0.0 ? nviol : 0;
}

// ====================================================================== compute warps
template <int CT, int LCH>
__device__ __forceinline__ void fast_compute_role(const FuseArgs& A) {
    // (pointers are derived from the shared array here so that the accesses compile to LDS/STS, not generic loads)
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int cap2 = (A.cap + 3) & ~1;                      // even => every sub-buffer stays 16-byte aligned
    double* const ts_s = reinterpret_cast<double*>(smem_raw);
    double* const pos_s = ts_s + cap2;
    double* const z_s = pos_s + 3 * (size_t)cap2;
    double* const sd = z_s + 3 * (size_t)cap2;
    int* const iscr = reinterpret_cast<int*>(sd + FS_INT);
    uint64_t* const mbar = reinterpret_cast<uint64_t*>(sd + FS_MBAR);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int NW = CT / 32;
    long long* const pclk = (A.phase_clock && blockIdx.x == 0 && lane == 0 && (tid == 0 || warp >= NW)) ? A.phase_clock : nullptr;
    (void)ts_s; (void)pos_s; (void)z_s; (void)iscr; (void)warp; (void)NW; (void)pclk;
    GSF_FSTAMP_DECL;

    uint32_t par_full = 0, par_q0 = 0, par_q1 = 0;
#ifdef GSF_DEBUG_STAMPS
    const long long blk_t0 = clock64();
#endif
    // the block's work queue: entries 0-2 were filled before the kernel's first barrier; entry j + 3 is requested at
    // the start of trajectory j and published at its end (before the slot is handed back to the look-ahead warps)
    TrajRef* const ring = reinterpret_cast<TrajRef*>(sd + FS_RING);
    TrajRef cur = ring[0];
    if (tid == 0 && cur.b < A.B) issue_trajectory_load_hint(A, cur.e0, cur.n, ts_s, pos_s, z_s, mbar + MB_FULL);
    int j = 0;
    for (;;) {
        if (cur.b >= A.B) break;
        const int b = cur.b, n = cur.n;
        const long long e0 = cur.e0;
        const TrajRef nxt = ring[(j + 1) & 3];
        cur = nxt;
        const bool has_next = nxt.b < A.B;
        const long long o0 = nxt.e0;
        const int n_next = nxt.n;
        int fetched = 0;
        if (tid == 0) fetched = atomicAdd(A.work_counter, 1);      // trajectory j + 3 (consumed after pass B)
        TrajRef* const ring_out = ring + ((j + 3) & 3);
        const int slot = j & 1;
        ++j;
        GSF_FSTAMP(0);
        double* bc = sd + FS_BC + 48 * slot;
        const double* pst = sd + FS_PST + 3 * CT * slot;
        const int lead = (int)(e0 & 1);
        double* tsS = ts_s + lead; double* posS = pos_s + 3 * lead; double* zS = z_s + 3 * lead;
        if (tid < 23 && (A.params_per_traj || j == 1))      // a batch-wide record is fetched once
            sd[FS_PRM + tid] = reinterpret_cast<const double*>(A.params + (A.params_per_traj ? b : 0))[tid];
        if (tid == 32 % CT) iscr[0] = 0;
        {
            const int cnt = n + lead, even = cnt & ~1;
            if ((cnt & 1) && tid < 7) {                   // odd tail element: plain copy
                const long long g = e0 - lead + even;
                if (tid == 0) ts_s[even] = A.ts[g];
                else if (tid < 4) pos_s[3 * even + (tid - 1)] = A.pos[3 * g + (tid - 1)];
                else z_s[3 * even + (tid - 4)] = A.z[3 * g + (tid - 4)];
            }
            if (even > 0) { mbar_wait_polite(mbar + MB_FULL, par_full); par_full ^= 1; }
        }
        GSF_FSTAMP(1);
        named_sync(NB_AUXRDY + slot, CT + 32);             // look-ahead results of this trajectory are published (and, as every
        __threadfence_block();                             // compute thread takes part, the parameter / tail writes above)
        GSF_FSTAMP(2);
        if (bc[30] != 0.0) {
            // needs the general machinery: leave it to the general kernel
            if (tid == 0) {
                A.status[b] = ST_DEFERRED;
                atomicAdd(A.defer_count, 1);
                if (has_next) issue_trajectory_load_hint(A, o0, n_next, ts_s, pos_s, z_s, mbar + MB_FULL);
                *ring_out = traj_ref_from(A, fetched);
                __threadfence_block();
            }
            named_sync(1, CT);                           // every thread has read the verdict
            if (GSF_A_EARLY && CT > 32) named_arrive(NB_MID + slot, CT + 32);
            named_arrive(NB_FREE_A + slot, CT + 32); named_arrive(NB_FREE_B + slot, CT + 32);
            continue;
        }

        // quaternions of this trajectory into L2 now; the streaming pass reads them ~10k cycles later
        if (tid == 0) { const long long qn = ((long long)n * 32) & ~15ll; bulk_prefetch_l2_hint(A.quat + 4 * e0, (uint32_t)qn, l2_policy_evict_last()); }
        const FuseParams& prm = *reinterpret_cast<const FuseParams*>(sd + FS_PRM);
        const bool xy_same = prm.p0[0] == prm.p0[1] && prm.q[0] == prm.q[1] && prm.r[0] == prm.r[1];
        const int c0 = min(tid * LCH, n), c1 = min(c0 + LCH, n);
        const int s0 = max(c0, 1);                      // steps owned: i in [s0, c1)
        int st = (int)bc[31];

        // ------------------------------------------------------------------ pass B: gains, affine maps, residual check
        const double thr2 = !(prm.residual_thresh > 0.0) ? -1.0 : prm.residual_thresh * prm.residual_thresh;
        double pprev0 = 0.0, pprev1 = 0.0, pprev2 = 0.0;
        if (c0 < n) { const int ip = max(c0 - 1, 0); pprev0 = posS[3 * ip]; pprev1 = posS[3 * ip + 1]; pprev2 = posS[3 * ip + 2]; }
        int nviol = 0;
        if (c0 == 0 && thr2 > 0.0) {                      // pose 0 against its Sim3 image
            double rx, ry, rz;
            mat_vec(bc, posS[0], posS[1], posS[2], rx, ry, rz);
            const double d0 = bc[19] * rx + bc[16] - zS[0], d1 = bc[19] * ry + bc[17] - zS[1], d2 = bc[19] * rz + bc[18] - zS[2];
            if (!(d0 * d0 + d1 * d1 + d2 * d2 < thr2)) ++nviol;
        }
        named_sync(1, CT);                               // neighbours' boundary poses are read before being overwritten
        Aff3 aff;
        if (xy_same) nviol += pass_b12<true>(tsS, posS, zS, bc, prm, pst + 3 * tid, s0, c1, thr2, pprev0, pprev1, pprev2, aff);
        else nviol += pass_b12<false>(tsS, posS, zS, bc, prm, pst + 3 * tid, s0, c1, thr2, pprev0, pprev1, pprev2, aff);
        if (thr2 > 0.0) {
            nviol = warp_sum_i(nviol);
            if (lane == 0 && nviol) atomicAdd(iscr, nviol);
        }
        Aff3 aex;                                       // exclusive affine prefix inside the warp
        aff_warp_scan(aff, lane);
        if (NW > 1 && lane == 31) {
#pragma unroll
            for (int k = 0; k < 3; ++k) { sd[FS_AFF + warp * 6 + k] = aff.a[k]; sd[FS_AFF + warp * 6 + 3 + k] = aff.b[k]; }
        }
#pragma unroll
        for (int k = 0; k < 3; ++k) { aex.a[k] = __shfl_up_sync(GSF_FULL_MASK, aff.a[k], 1); aex.b[k] = __shfl_up_sync(GSF_FULL_MASK, aff.b[k], 1); }
        if (lane == 0) {
#pragma unroll
            for (int k = 0; k < 3; ++k) { aex.a[k] = 1.0; aex.b[k] = 0.0; }
        }
        fence_proxy_async();                            // generic reads of the timestamp buffer before the TMA write below
        named_sync(1, CT);
        // Quaternions, part 1: the timestamp buffer is dead after pass B, so the first nq1 quaternions (32 B each; nq1 a
        // multiple of 32 poses) land in it while pass C runs; part 2 follows into the position buffer after pass C.
        const int nq1 = min(n, (cap2 >> 2) & ~31);
        if (tid == 0) {
            mbar_expect_tx(mbar + MB_QUAT0, (uint32_t)nq1 * 32u);
            if (nq1 > 0) bulk_g2s_hint(ts_s, A.quat + 4 * e0, (uint32_t)nq1 * 32u, mbar + MB_QUAT0, l2_policy_evict_first());
        }
        TrajRef fref{0, 0, 0};
        if (tid == 0) fref = traj_ref_from(A, fetched);
        GSF_FSTAMP(3);
        if (GSF_A_EARLY == 1 && CT > 32) named_arrive(NB_MID + slot, CT + 32);     // warp A may start on trajectory j + 2

        // ------------------------------------------------------------------ pass C: state recursion
        {
            Aff3 pre = aex;
            if (NW > 1 && warp > 0) {
                Aff3 acc;
#pragma unroll
                for (int k = 0; k < 3; ++k) { acc.a[k] = sd[FS_AFF + k]; acc.b[k] = sd[FS_AFF + 3 + k]; }
                for (int w = 1; w < warp; ++w) {
                    Aff3 nx;
#pragma unroll
                    for (int k = 0; k < 3; ++k) { nx.a[k] = sd[FS_AFF + w * 6 + k]; nx.b[k] = sd[FS_AFF + w * 6 + 3 + k]; }
                    acc = aff_compose(acc, nx);
                }
                pre = aff_compose(acc, aex);
            }
            double x0 = pre.a[0] * bc[13] + pre.b[0], x1 = pre.a[1] * bc[14] + pre.b[1], x2 = pre.a[2] * bc[15] + pre.b[2];
            if (c0 == 0) { zS[0] = bc[13]; zS[1] = bc[14]; zS[2] = bc[15]; }
#pragma unroll PASSC_UNROLL
            for (int i = s0; i < c1; ++i) {
                x0 = posS[3 * i] * x0 + zS[3 * i]; x1 = posS[3 * i + 1] * x1 + zS[3 * i + 1]; x2 = posS[3 * i + 2] * x2 + zS[3 * i + 2];
                zS[3 * i] = x0; zS[3 * i + 1] = x1; zS[3 * i + 2] = x2;
            }
        }
        fence_proxy_async();
        named_sync(1, CT);
        GSF_FSTAMP(4);
        if (GSF_A_EARLY == 2 && CT > 32) named_arrive(NB_MID + slot, CT + 32);

        // ------------------------------------------------------------------ store fused positions; stream the quaternions
        double* gout = A.out_pos + 3 * e0;
        const int viol_total = iscr[0];
        if (viol_total) st |= ST_RANSAC_OUTLIERS;
        if (tid == 0) {
            const int m = n - lead, even = m & ~1;
            if (even > 0) { bulk_s2g_hint(gout + 3 * lead, zS + 3 * lead, (uint32_t)even * 24u, l2_policy_evict_first()); bulk_commit(); }
            if (lead) { gout[0] = zS[0]; gout[1] = zS[1]; gout[2] = zS[2]; }
            if (m & 1) {
                const int q = 3 * (lead + even);
                gout[q] = zS[q]; gout[q + 1] = zS[q + 1]; gout[q + 2] = zS[q + 2];
            }
            A.status[b] = st;
            if (A.sim3_out) {
                double* o = A.sim3_out + 16 * (size_t)b;
#pragma unroll
                for (int k = 0; k < 9; ++k) o[k] = bc[20 + k];
                o[9] = bc[16]; o[10] = bc[17]; o[11] = bc[18]; o[12] = bc[19];
                o[13] = (double)n; o[14] = (double)n; o[15] = (double)viol_total;
            }
        }
        {
            // Quaternions q_state[i] = C (x) q_hat[i]: the timestamp + position buffers (32 B/pose) are dead after
            // pass C, so the whole quaternion array comes in with ONE TMA bulk copy (it was prefetched into L2 at the
            // start of the trajectory) and the per-pose loop is ~60 instructions -- the unrolled global->global
            // version was ~600, which thrashes the ~8 KB per-partition instruction cache (tools/micro/icache_roles.cu).
            double* qs = ts_s;                                  // [n,4] over ts_s + pos_s
            if (tid == 0 && n > nq1) {
                mbar_expect_tx(mbar + MB_QUAT, (uint32_t)(n - nq1) * 32u);
                bulk_g2s_hint(qs + 4 * nq1, A.quat + 4 * (e0 + nq1), (uint32_t)(n - nq1) * 32u, mbar + MB_QUAT, l2_policy_evict_first());
            }
            const Quat C{bc[9], bc[10], bc[11], bc[12]};
            const uint64_t pf = l2_policy_evict_first();
            double2* __restrict__ qout = reinterpret_cast<double2*>(A.out_quat + 4 * e0);
            const double2* __restrict__ q2 = reinterpret_cast<const double2*>(qs);
            mbar_wait_polite(mbar + MB_QUAT0, par_q0); par_q0 ^= 1;
            int bad = 0;
            // one pose per iteration, branch-free reciprocal square root (two interleaved poses per iteration measured
            // 2 % slower with the roles inlined: instruction-cache footprint again).  Two phases over the same loop
            // body: warp-rows that lie entirely in part 1, then -- after part 2 has landed -- the rest.
            int i0 = tid - lane;                                // warp-uniform loop variables (the wait below syncs the warp)
            constexpr int QS = (GSF_QUAT2 + 1) * CT;            // poses per block and iteration
            int stop = min(n, nq1) - (QS - CT);                 // nq1 is a multiple of 32: phase 0 takes the iterations whose rows all lie in part 1
#pragma unroll 1
            for (int phase = 0; phase < 2; ++phase) {
#pragma unroll 1
                for (; i0 < stop; i0 += QS) {
                    Quat rr[GSF_QUAT2 + 1];
#pragma unroll
                    for (int u = 0; u <= GSF_QUAT2; ++u) {
                        const int i = i0 + u * CT + lane;
                        const int ia = i < n ? i : 0;           // the tail re-reads a landed pose and stores nothing
                        const double2 lo0 = q2[2 * ia], hi0 = q2[2 * ia + 1];
                        const Quat qa{lo0.x, lo0.y, hi0.x, hi0.y};
                        const double na = qnorm2(qa);
                        if (na == 0.0) bad = 1;                 // scipy raises here (:466); output row becomes NaN
                        rr[u] = qscale(qmul(C, qa), fast_rsqrt(na));
                    }
#pragma unroll
                    for (int u = 0; u <= GSF_QUAT2; ++u) {
                        const int i = i0 + u * CT + lane;
                        if (i < n) {
                            stg2_hint(qout + 2 * i, make_double2(rr[u].x, rr[u].y), pf);
                            stg2_hint(qout + 2 * i + 1, make_double2(rr[u].z, rr[u].w), pf);
                        }
                    }
                }
                if (phase == 0 && n > nq1) { mbar_wait_polite(mbar + MB_QUAT, par_q1); par_q1 ^= 1; }
                stop = n;
            }
            GSF_FSTAMP(5);
            fence_proxy_async();                                // generic reads of the buffer before the next TMA writes
            named_sync(1, CT);                                  // status[b] is written; every thread is done with the buffers
            if (tid == 0) {
                bulk_wait_read();                               // the position store has left shared memory
                fence_proxy_async();
                if (has_next) issue_trajectory_load_hint(A, o0, n_next, ts_s, pos_s, z_s, mbar + MB_FULL);
                *ring_out = fref;
                __threadfence_block();
            }
            named_arrive(NB_FREE_A + slot, CT + 32); named_arrive(NB_FREE_B + slot, CT + 32);
            if (bad) atomicOr(A.status + b, ST_BAD_QUATERNION);
            GSF_FSTAMP(6);
        }
    }
#ifdef GSF_DEBUG_STAMPS
    if (A.phase_clock && tid == 0 && blockIdx.x % 37 == 0 && blockIdx.x / 37 < 12) {      // debug: per-block totals
        long long* o = A.phase_clock + 64 + 4 * (blockIdx.x / 37);
        unsigned smid; asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
        o[0] = clock64() - blk_t0; o[1] = j; o[2] = smid; o[3] = blk_t0;
    }
#endif
}

// Pivot-shifted Umeyama sums of the pose pairs [pair_lo, pair_hi) of one trajectory, streamed from global memory by one
// warp (lane-strided pairs, NP pairs in flight per lane, fixed order -> bit-reproducible).
template <int NP, bool LDG = true>
__device__ __forceinline__ void stream_pose_sums(const double* __restrict__ gp, const double* __restrict__ gz, long long e0, int n,
                                                 int pair_lo, int pair_hi, int lane, double* v,
                                                 double ps0, double ps1, double ps2, double pz0, double pz1, double pz2) {
    if (!(e0 & 1)) {
        const double2* __restrict__ gp2 = reinterpret_cast<const double2*>(gp);
        const double2* __restrict__ gz2 = reinterpret_cast<const double2*>(gz);
#pragma unroll 1
        for (int p = pair_lo + lane; p < pair_hi; p += 32 * NP) {
            // NP pose pairs per round: 6 NP 128-bit loads in flight
            double2 a[NP][3], c[NP][3];
#pragma unroll
            for (int h = 0; h < NP; ++h) {
                const int pp = p + 32 * h;
                if (pp < pair_hi && 2 * pp + 1 < n) {
                    // plain read-only loads: the lines were requested with evict_last by the bulk prefetch, and a per-load
                    // cache-policy operand costs two uniform-register moves per load in this loop
#pragma unroll
                    for (int q = 0; q < 3; ++q) { a[h][q] = LDG ? __ldg(gp2 + 3 * pp + q) : gp2[3 * pp + q]; c[h][q] = LDG ? __ldg(gz2 + 3 * pp + q) : gz2[3 * pp + q]; }
                } else if (pp < pair_hi && 2 * pp < n) {        // last pose of an odd-length trajectory
                    a[h][0] = make_double2(gp[6 * pp], gp[6 * pp + 1]); a[h][1] = make_double2(gp[6 * pp + 2], 0.0);
                    c[h][0] = make_double2(gz[6 * pp], gz[6 * pp + 1]); c[h][1] = make_double2(gz[6 * pp + 2], 0.0);
                    a[h][2] = make_double2(0.0, 0.0); c[h][2] = make_double2(0.0, 0.0);
                }
            }
#pragma unroll
            for (int h = 0; h < NP; ++h) {
                const int pp = p + 32 * h;
                if (pp < pair_hi && 2 * pp < n) umeyama_accumulate(v, a[h][0].x, a[h][0].y, a[h][1].x, c[h][0].x, c[h][0].y, c[h][1].x, ps0, ps1, ps2, pz0, pz1, pz2);
                if (pp < pair_hi && 2 * pp + 1 < n) umeyama_accumulate(v, a[h][1].y, a[h][2].x, a[h][2].y, c[h][1].y, c[h][2].x, c[h][2].y, ps0, ps1, ps2, pz0, pz1, pz2);
            }
        }
    } else {
        // odd pose offset: rows are only 8-byte aligned; same pose order, 64-bit loads
#pragma unroll 1
        for (int p = pair_lo + lane; p < pair_hi; p += 32) {
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int i = 2 * p + h;
                if (i < n) umeyama_accumulate(v, gp[3 * i], gp[3 * i + 1], gp[3 * i + 2], gz[3 * i], gz[3 * i + 1], gz[3 * i + 2], ps0, ps1, ps2, pz0, pz1, pz2);
            }
        }
    }
}
// Pose pairs (the LAST ones of the trajectory, whole 64-pair rounds) that the scan warp sums itself while it would otherwise
// wait for the sums warp: the look-ahead chain is sums -> SVD, and the sums are its longer half (14.8 k of 23.6 k cycles at
// 1000 poses) -- with a quarter of them on warp B, which idles ~6 k cycles per trajectory between its covariance scan and
// the SVD, both warps finish together.  Long trajectories only (two compute warps).  Measured (round 2): warp A's streaming
// drops from 14.8 k to 11.8 k cycles and the chain gets ~1 k cycles of slack, but the block's period stays at 23 k (31.5 vs
// 30.8 ms per 2^20 x 1000 poses): the compute warps are as long, and every phase stretches when another shrinks -- the SM's
// FP64 issue capacity is what the three blocks share.  Off by default.
template <int CT>
__device__ __forceinline__ int sums_share_b(int npairs) {
    return (GSF_SUMS_SPLIT && CT > 32) ? ((npairs / 4 + 32) & ~63) : 0;
}

template <int CT, int LCH>
__device__ __forceinline__ void fast_sums_role(const FuseArgs& A) {
    // (pointers are derived from the shared array here so that the accesses compile to LDS/STS, not generic loads)
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int cap2 = (A.cap + 3) & ~1;                      // even => every sub-buffer stays 16-byte aligned
    double* const ts_s = reinterpret_cast<double*>(smem_raw);
    double* const pos_s = ts_s + cap2;
    double* const z_s = pos_s + 3 * (size_t)cap2;
    double* const sd = z_s + 3 * (size_t)cap2;
    int* const iscr = reinterpret_cast<int*>(sd + FS_INT);
    uint64_t* const mbar = reinterpret_cast<uint64_t*>(sd + FS_MBAR);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int NW = CT / 32;
    long long* const pclk = (A.phase_clock && blockIdx.x == 0 && lane == 0 && (tid == 0 || warp >= NW)) ? A.phase_clock : nullptr;
    (void)ts_s; (void)pos_s; (void)z_s; (void)iscr; (void)warp; (void)NW; (void)pclk;
    GSF_FSTAMP_DECL;

    // ====================================================================== warp A: Umeyama sums, one to two trajectories ahead
    const TrajRef* const ring = reinterpret_cast<const TrajRef*>(sd + FS_RING);
    int j = 0;
    for (;;) {
        const int slot = j & 1, k = j >> 1;
        GSF_FSTAMP(16);
        // Short trajectories (CT = 32: up to 288 poses, 5 blocks per SM): the streaming runs ahead of the slot -- the
        // slot is only needed for the hand-over below -- which measured 5 % faster at 271 poses; at 1000 poses the
        // extra trajectory of look-ahead falls out of L2 (DRAM reads 5.9 -> 8.3 GB per 65536 trajectories, 3 % slower).
        constexpr bool AHEAD = CT <= 32;
        constexpr bool EARLY = !AHEAD && GSF_A_EARLY != 0;   // start when trajectory j - 2 is past pass B / C; the slot itself is only needed for the hand-over
        if (EARLY) { if (k > 0) named_sync(NB_MID + slot, CT + 32); }
        else if (!AHEAD && k > 0) named_sync(NB_FREE_A + slot, CT + 32);     // also publishes the queue entry of this trajectory
        const TrajRef ref = ring[j & 3];
        if (ref.b >= A.B) break;
        const long long e0 = ref.e0;
        const int n = ref.n;
        ++j;
        GSF_FSTAMP(17);
        if (lane == 0) prefetch_pos_z(A, e0, n);          // whole trajectory into L2 now: rounds after the first hit L2
        const double* __restrict__ gp = A.pos + 3 * e0;
        const double* __restrict__ gz = A.z + 3 * e0;
        const double ps0 = gp[0], ps1 = gp[1], ps2 = gp[2], pz0 = gz[0], pz1 = gz[1], pz2 = gz[2];
        double v[16];
#pragma unroll
        for (int q = 0; q < 16; ++q) v[q] = 0.0;
        constexpr int NP = CT <= 32 ? 1 : 2;                // pose pairs in flight per lane (short trajectories, 5 blocks per SM: one measured 4 % faster)
        const int npairs = (n + 1) >> 1;
        stream_pose_sums<NP>(gp, gz, e0, n, 0, npairs - sums_share_b<CT>(npairs), lane, v, ps0, ps1, ps2, pz0, pz1, pz2);
        GSF_FSTAMP(18);
        const double total = butterfly16(v, lane);
        double* sums = sd + FS_SUMS + 24 * slot;
        // (AHEAD) the slot is only needed now; this wait also publishes the queue entries up to trajectory j + 2,
        // read at the top of the next iterations
        if ((AHEAD || EARLY) && k > 0) named_sync(NB_FREE_A + slot, CT + 32);
        if (!(lane & 1)) sums[butterfly16_index(lane)] = total;
        if (lane == 1) { sums[16] = ps0; sums[17] = ps1; sums[18] = ps2; sums[19] = pz0; sums[20] = pz1; sums[21] = pz2; }
        __threadfence_block();
        named_arrive(NB_SUMS + slot, 64);
        GSF_FSTAMP(19);
    }
}

template <int CT, int LCH>
__device__ __forceinline__ void fast_scan_svd_role(const FuseArgs& A) {
    // (pointers are derived from the shared array here so that the accesses compile to LDS/STS, not generic loads)
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int cap2 = (A.cap + 3) & ~1;                      // even => every sub-buffer stays 16-byte aligned
    double* const ts_s = reinterpret_cast<double*>(smem_raw);
    double* const pos_s = ts_s + cap2;
    double* const z_s = pos_s + 3 * (size_t)cap2;
    double* const sd = z_s + 3 * (size_t)cap2;
    int* const iscr = reinterpret_cast<int*>(sd + FS_INT);
    uint64_t* const mbar = reinterpret_cast<uint64_t*>(sd + FS_MBAR);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int NW = CT / 32;
    long long* const pclk = (A.phase_clock && blockIdx.x == 0 && lane == 0 && (tid == 0 || warp >= NW)) ? A.phase_clock : nullptr;
    (void)ts_s; (void)pos_s; (void)z_s; (void)iscr; (void)warp; (void)NW; (void)pclk;
    GSF_FSTAMP_DECL;

    // ====================================================================== warp B: covariance start values, gap/window check, Umeyama finish
    double* tsb = sd + FS_PST + 6 * CT;
    uint32_t par_ts = 0;
    const TrajRef* const ring = reinterpret_cast<const TrajRef*>(sd + FS_RING);
    if (lane == 0 && ring[0].b < A.B) issue_ts_load(A, ring[0].e0, ring[0].n, tsb, mbar + MB_TSB);
    int j = 0;
    for (;;) {
        const int slot = j & 1, k = j >> 1;
        GSF_FSTAMP(24);
        if (k > 0) named_sync(NB_FREE_B + slot, CT + 32);     // also publishes the queue entries up to trajectory j + 1
        const TrajRef ref = ring[j & 3];
        if (ref.b >= A.B) break;
        const int b = ref.b, n = ref.n;
        const long long e0 = ref.e0;
        ++j;
        // the global scalars this warp needs are requested here and consumed after the scan: the first quaternion
        // (lanes 0-3 hold one component each), the next trajectory's offsets (above) and, one trajectory ahead, the parameters
        const double q0c = A.quat[4 * e0 + (lane & 3)];
        GSF_FSTAMP(25);
        // parameters: this warp's shared-memory copy (a batch-wide record is fetched once)
        if (A.params_per_traj || k + slot == 0) {
            if (lane < 23) sd[FS_PRMB + lane] = reinterpret_cast<const double*>(A.params + (A.params_per_traj ? b : 0))[lane];
            __syncwarp();
        }
        const FuseParams* __restrict__ gprm = reinterpret_cast<const FuseParams*>(sd + FS_PRMB);
        const bool xy_same = gprm->p0[0] == gprm->p0[1] && gprm->q[0] == gprm->q[1] && gprm->r[0] == gprm->r[1];
        double* pst = sd + FS_PST + 3 * CT * slot;
        {
            const int lead = (int)(e0 & 1), cnt = n + lead, even = cnt & ~1;
            if ((cnt & 1) && lane == 0) tsb[even] = A.ts[e0 - lead + even];      // odd tail element: plain copy
            mbar_wait_polite(mbar + MB_TSB, par_ts); par_ts ^= 1;
            __syncwarp();
        }
#ifdef GSF_DEBUG_STAMPS
        long long* clk = (A.phase_clock && blockIdx.x == 0 && lane == 0 && j == 100) ? A.phase_clock : nullptr;
        if (clk) clk[31] = clock64();
#else
        long long* const clk = nullptr;
#endif
        int general = xy_same ? cov_start_scan<2, CT, LCH>(tsb + (e0 & 1), n, gprm, lane, pst, clk)
                              : cov_start_scan<3, CT, LCH>(tsb + (e0 & 1), n, gprm, lane, pst, clk);
        __syncwarp();
        if (lane == 0) {                                    // next trajectory's timestamps (its queue entry was published with this slot's release)
            const TrajRef nx = ring[j & 3];
            if (nx.b < A.B) issue_ts_load(A, nx.e0, nx.n, tsb, mbar + MB_TSB);
        }
        GSF_FSTAMP(26);
        // this warp's share of the Umeyama sums (the last pose pairs), while warp A streams the rest
        double* const sums_b = sd + FS_PST + 6 * CT + cap2;
        const int npairs = (n + 1) >> 1, share = sums_share_b<CT>(npairs);
        if (share > 0) {
            const double* __restrict__ gp = A.pos + 3 * e0;
            const double* __restrict__ gz = A.z + 3 * e0;
            double vb[16];
#pragma unroll
            for (int q = 0; q < 16; ++q) vb[q] = 0.0;
            stream_pose_sums<2>(gp, gz, e0, n, npairs - share, npairs, lane, vb, gp[0], gp[1], gp[2], gz[0], gz[1], gz[2]);
            const double tb = butterfly16(vb, lane);
            if (!(lane & 1)) sums_b[butterfly16_index(lane)] = tb;
            __syncwarp();
        }
        named_sync(NB_SUMS + slot, 64);
        __threadfence_block();
        GSF_FSTAMP(27);
        const double* sums = sd + FS_SUMS + 24 * slot;
        double v[16], chk = 0.0;
#pragma unroll
        for (int q = 0; q < 16; ++q) { v[q] = share > 0 ? sums[q] + sums_b[q] : sums[q]; chk += v[q]; }
        if (!(fabs(chk) <= 1.7976931348623157e308)) general = 1;       // NaN row (no GNSS) or non-finite input
        if (n < 3 || n < gprm->min_samples) general = 1;
        const Quat q0{__shfl_sync(GSF_FULL_MASK, q0c, 0), __shfl_sync(GSF_FULL_MASK, q0c, 1), __shfl_sync(GSF_FULL_MASK, q0c, 2),
                      __shfl_sync(GSF_FULL_MASK, q0c, 3)};
        if (qnorm2(q0) == 0.0) general = 1;
        int ust = 0;
        double* bc = sd + FS_BC + 48 * slot;
        if (!general) {
            const double nn = (double)n, inv = 1.0 / nn;
            const double ma[3] = {v[0] * inv, v[1] * inv, v[2] * inv}, mb[3] = {v[3] * inv, v[4] * inv, v[5] * inv};
            double ms_[3], md_[3], hh[9];
#pragma unroll
            for (int q = 0; q < 3; ++q) { ms_[q] = sums[16 + q] + ma[q]; md_[q] = sums[19 + q] + mb[q]; }
#pragma unroll
            for (int r = 0; r < 3; ++r)
#pragma unroll
                for (int c = 0; c < 3; ++c) hh[3 * r + c] = v[6 + 3 * r + c] - nn * ma[r] * mb[c];
            const double ss = v[15] - nn * (ma[0] * ma[0] + ma[1] * ma[1] + ma[2] * ma[2]);
            double R[9], t[3], s = 1.0;
            ust = umeyama_finish(n, ms_, md_, hh, ss, R, t, s);
            if (lane == 0) {
                const Quat qR = quat_from_matrix(R);
                const Quat q0h = qunit(q0);
                const Quat qs0 = qunit_or_identity(qmul(qR, q0h));
                const Quat Cq = qmul(qs0, qconj(q0h));
                double M[9]; qmat(Cq, M);
#pragma unroll
                for (int q = 0; q < 9; ++q) { bc[q] = M[q]; bc[20 + q] = R[q]; }
                bc[9] = Cq.x; bc[10] = Cq.y; bc[11] = Cq.z; bc[12] = Cq.w;
                double rx, ry, rz;
                mat_vec(R, sums[16], sums[17], sums[18], rx, ry, rz);
                bc[13] = s * rx + t[0]; bc[14] = s * ry + t[1]; bc[15] = s * rz + t[2];
                bc[16] = t[0]; bc[17] = t[1]; bc[18] = t[2]; bc[19] = s;
            }
        }
        if (lane == 0) { bc[30] = general ? 1.0 : 0.0; bc[31] = (double)ust; }
        __threadfence_block();
        named_arrive(NB_AUXRDY + slot, CT + 32);
        GSF_FSTAMP(28);
    }
}

template <int CT, int LCH>
__global__ void __launch_bounds__(CT + 64, fast_min_blocks(CT)) fuse_fast_kernel(const __grid_constant__ FuseArgs A) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    constexpr int NW = CT / 32;
    const int cap2 = (A.cap + 3) & ~1;
    uint64_t* mbar = reinterpret_cast<uint64_t*>(reinterpret_cast<double*>(smem_raw) + 7 * (size_t)cap2 + FS_MBAR);
    const int warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) {
#pragma unroll
        for (int k = 0; k < 8; ++k) mbar_init(mbar + k, 1);
        fence_mbar_init();
        TrajRef* ring = reinterpret_cast<TrajRef*>(reinterpret_cast<double*>(smem_raw) + 7 * (size_t)cap2 + FS_RING);
        // one after the other: an entry that skips an unusable trajectory draws a new index, and that index must be
        // larger than those of the entries before it and smaller than those after it -- the end-of-sequence entry
        // (b >= B) has to be the last one of a block
        for (int k = 0; k < 3; ++k) ring[k] = traj_ref_from(A, atomicAdd(A.work_counter, 1));
    }
    __syncthreads();
    if (warp < NW) fast_compute_role<CT, LCH>(A);
    else if (warp == NW) fast_sums_role<CT, LCH>(A);
    else fast_scan_svd_role<CT, LCH>(A);
}

// ====================================================================== short trajectories: one warp per trajectory
// Up to 32 * LCH poses.  The three roles above pay per trajectory for their hand-offs, for the second read of positions and
// measurements and for a two-stage pipeline that has to fill -- fixed costs that a 271-pose trajectory does not amortise
// (4096 x 271 poses: 31 % of the HBM roofline with the role kernel, its SVD warp busy 13 of every 15.8 k cycles).  Here a
// block is ONE warp that takes a trajectory through the same steps by itself -- TMA load, covariance start scan, Umeyama
// sums (from shared memory), SVD, pass B, warp scan, pass C, bulk store, quaternions -- with the same device functions in
// the same order (bit-identical results), and the SM overlaps a dozen such warps at different stages instead of three
// roles in lock step.  Work is handed out by the global counter.
// MEASURED (round 2, tools/warp_vs_roles.py, profiles/r02_warp_kernel_*): bit-identical to the role kernel and 1.7x SLOWER
// (262 144 x 271 poses: 5.36 vs 3.13 ms; 4096 x 271: 0.149 vs 0.102 ms).  Every warp walks the whole 72 KB of code
// (4536 instructions) by itself, twelve warps per SM at twelve different places: instruction-cache hit rate 72 %, and
// `no_instruction` + `branch_resolving` (the rolled shuffle scans) are the top stalls at 33 % issue utilisation -- the
// role split is what keeps each warp inside a small loop.  Kept behind GSF_FAST_CT=1 (test_warp_kernel_matches_role_kernel).
constexpr int WS_BC = 0, WS_SUMS = 48, WS_PRM = 72, WS_PST = 96, WS_MBAR = 192, WS_TOTAL = 200;      // doubles after the staging buffers
__host__ __device__ constexpr size_t warp_smem_bytes(int cap) { return (size_t)((cap + 3) & ~1) * 56 + (size_t)WS_TOTAL * 8; }

template <int LCH>
__global__ void __launch_bounds__(32, 12) fuse_warp_kernel(const __grid_constant__ FuseArgs A) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int cap2 = (A.cap + 3) & ~1;
    double* const ts_s = reinterpret_cast<double*>(smem_raw);
    double* const pos_s = ts_s + cap2;
    double* const z_s = pos_s + 3 * (size_t)cap2;
    double* const sd = z_s + 3 * (size_t)cap2;
    uint64_t* const mbar = reinterpret_cast<uint64_t*>(sd + WS_MBAR);      // 0 trajectory, 1 quaternions part 1, 2 part 2
    const int lane = threadIdx.x;
    if (lane == 0) {
        mbar_init(mbar, 1); mbar_init(mbar + 1, 1); mbar_init(mbar + 2, 1);
        fence_mbar_init();
    }
    __syncwarp();
    uint32_t par_full = 0, par_q0 = 0, par_q1 = 0;
    bool have_prm = false;
    for (;;) {
        // ---- next trajectory
        TrajRef ref{0, 0, 0};
        if (lane == 0) ref = traj_ref_from(A, atomicAdd(A.work_counter, 1));
        const int b = __shfl_sync(GSF_FULL_MASK, ref.b, 0), n = __shfl_sync(GSF_FULL_MASK, ref.n, 0);
        const long long e0 = __shfl_sync(GSF_FULL_MASK, ref.e0, 0);
        if (b >= A.B) break;
        const int lead = (int)(e0 & 1);
        if (lane == 0) {
            issue_trajectory_load_hint(A, e0, n, ts_s, pos_s, z_s, mbar);
            const long long qn = ((long long)n * 32) & ~15ll;
            bulk_prefetch_l2_hint(A.quat + 4 * e0, (uint32_t)qn, l2_policy_evict_last());
        }
        const double q0c = A.quat[4 * e0 + (lane & 3)];
        if (lane < 23 && (A.params_per_traj || !have_prm))      // a batch-wide record is fetched once
            sd[WS_PRM + lane] = reinterpret_cast<const double*>(A.params + (A.params_per_traj ? b : 0))[lane];
        have_prm = true;
        double* tsS = ts_s + lead; double* posS = pos_s + 3 * lead; double* zS = z_s + 3 * lead;
        {
            const int cnt = n + lead, even = cnt & ~1;
            if ((cnt & 1) && lane < 7) {                  // odd tail element: plain copy
                const long long g = e0 - lead + even;
                if (lane == 0) ts_s[even] = A.ts[g];
                else if (lane < 4) pos_s[3 * even + (lane - 1)] = A.pos[3 * g + (lane - 1)];
                else z_s[3 * even + (lane - 4)] = A.z[3 * g + (lane - 4)];
            }
            if (even > 0) { mbar_wait_polite(mbar, par_full); par_full ^= 1; }
        }
        __syncwarp();
        const FuseParams& prm = *reinterpret_cast<const FuseParams*>(sd + WS_PRM);
        const bool xy_same = prm.p0[0] == prm.p0[1] && prm.q[0] == prm.q[1] && prm.r[0] == prm.r[1];
        double* const bc = sd + WS_BC;
        double* const pst = sd + WS_PST;
        double* const sums = sd + WS_SUMS;

        // ---- covariance start values + gap / window check (the scan warp's part 1)
        int general = xy_same ? cov_start_scan<2, 32, LCH>(tsS, n, &prm, lane, pst, nullptr)
                              : cov_start_scan<3, 32, LCH>(tsS, n, &prm, lane, pst, nullptr);
        // ---- Umeyama sums from the staged tile (the sums warp), finish + SVD (the scan warp's part 2)
        {
            const double ps0 = posS[0], ps1 = posS[1], ps2 = posS[2], pz0 = zS[0], pz1 = zS[1], pz2 = zS[2];
            double v[16];
#pragma unroll
            for (int q = 0; q < 16; ++q) v[q] = 0.0;
            stream_pose_sums<1, false>(posS, zS, e0, n, 0, (n + 1) >> 1, lane, v, ps0, ps1, ps2, pz0, pz1, pz2);
            const double total = butterfly16(v, lane);
            if (!(lane & 1)) sums[butterfly16_index(lane)] = total;
            if (lane == 1) { sums[16] = ps0; sums[17] = ps1; sums[18] = ps2; sums[19] = pz0; sums[20] = pz1; sums[21] = pz2; }
        }
        __syncwarp();
        int ust = 0;
        {
            double v[16], chk = 0.0;
#pragma unroll
            for (int q = 0; q < 16; ++q) { v[q] = sums[q]; chk += v[q]; }
            if (!(fabs(chk) <= 1.7976931348623157e308)) general = 1;       // NaN row (no GNSS) or non-finite input
            if (n < 3 || n < prm.min_samples) general = 1;
            const Quat q0{__shfl_sync(GSF_FULL_MASK, q0c, 0), __shfl_sync(GSF_FULL_MASK, q0c, 1), __shfl_sync(GSF_FULL_MASK, q0c, 2),
                          __shfl_sync(GSF_FULL_MASK, q0c, 3)};
            if (qnorm2(q0) == 0.0) general = 1;
            if (!general) {
                const double nn = (double)n, inv = 1.0 / nn;
                const double ma[3] = {v[0] * inv, v[1] * inv, v[2] * inv}, mb[3] = {v[3] * inv, v[4] * inv, v[5] * inv};
                double ms_[3], md_[3], hh[9];
#pragma unroll
                for (int q = 0; q < 3; ++q) { ms_[q] = sums[16 + q] + ma[q]; md_[q] = sums[19 + q] + mb[q]; }
#pragma unroll
                for (int r = 0; r < 3; ++r)
#pragma unroll
                    for (int c = 0; c < 3; ++c) hh[3 * r + c] = v[6 + 3 * r + c] - nn * ma[r] * mb[c];
                const double ss = v[15] - nn * (ma[0] * ma[0] + ma[1] * ma[1] + ma[2] * ma[2]);
                double R[9], t[3], s = 1.0;
                ust = umeyama_finish(n, ms_, md_, hh, ss, R, t, s);
                if (lane == 0) {
                    const Quat qR = quat_from_matrix(R);
                    const Quat q0h = qunit(q0);
                    const Quat qs0 = qunit_or_identity(qmul(qR, q0h));
                    const Quat Cq = qmul(qs0, qconj(q0h));
                    double M[9]; qmat(Cq, M);
#pragma unroll
                    for (int q = 0; q < 9; ++q) { bc[q] = M[q]; bc[20 + q] = R[q]; }
                    bc[9] = Cq.x; bc[10] = Cq.y; bc[11] = Cq.z; bc[12] = Cq.w;
                    double rx, ry, rz;
                    mat_vec(R, sums[16], sums[17], sums[18], rx, ry, rz);
                    bc[13] = s * rx + t[0]; bc[14] = s * ry + t[1]; bc[15] = s * rz + t[2];
                    bc[16] = t[0]; bc[17] = t[1]; bc[18] = t[2]; bc[19] = s;
                }
            }
        }
        __syncwarp();
        if (general) {
            // needs the general machinery: leave it to the general kernel
            if (lane == 0) { A.status[b] = ST_DEFERRED; atomicAdd(A.defer_count, 1); }
            fence_proxy_async();
            __syncwarp();
            continue;
        }

        // ---- pass B: gains, affine maps, residual check (the compute warps' code with one warp)
        const int c0 = min(lane * LCH, n), c1 = min(c0 + LCH, n);
        const int s0 = max(c0, 1);
        int st = ust;
        const double thr2 = !(prm.residual_thresh > 0.0) ? -1.0 : prm.residual_thresh * prm.residual_thresh;
        double pprev0 = 0.0, pprev1 = 0.0, pprev2 = 0.0;
        if (c0 < n) { const int ip = max(c0 - 1, 0); pprev0 = posS[3 * ip]; pprev1 = posS[3 * ip + 1]; pprev2 = posS[3 * ip + 2]; }
        int nviol = 0;
        if (c0 == 0 && thr2 > 0.0) {                      // pose 0 against its Sim3 image
            double rx, ry, rz;
            mat_vec(bc, posS[0], posS[1], posS[2], rx, ry, rz);
            const double d0 = bc[19] * rx + bc[16] - zS[0], d1 = bc[19] * ry + bc[17] - zS[1], d2 = bc[19] * rz + bc[18] - zS[2];
            if (!(d0 * d0 + d1 * d1 + d2 * d2 < thr2)) ++nviol;
        }
        __syncwarp();                                     // neighbours' boundary poses are read before being overwritten
        Aff3 aff;
        if (xy_same) nviol += pass_b12<true>(tsS, posS, zS, bc, prm, pst + 3 * lane, s0, c1, thr2, pprev0, pprev1, pprev2, aff);
        else nviol += pass_b12<false>(tsS, posS, zS, bc, prm, pst + 3 * lane, s0, c1, thr2, pprev0, pprev1, pprev2, aff);
        const int viol_total = thr2 > 0.0 ? warp_sum_i(nviol) : 0;
        Aff3 aex;                                         // exclusive affine prefix inside the warp
        aff_warp_scan(aff, lane);
#pragma unroll
        for (int k = 0; k < 3; ++k) { aex.a[k] = __shfl_up_sync(GSF_FULL_MASK, aff.a[k], 1); aex.b[k] = __shfl_up_sync(GSF_FULL_MASK, aff.b[k], 1); }
        if (lane == 0) {
#pragma unroll
            for (int k = 0; k < 3; ++k) { aex.a[k] = 1.0; aex.b[k] = 0.0; }
        }
        fence_proxy_async();                              // generic reads of the timestamp buffer before the TMA write below
        __syncwarp();
        // quaternions, part 1, into the dead timestamp buffer while pass C runs; part 2 into the position buffer after it
        const int nq1 = min(n, (cap2 >> 2) & ~31);
        if (lane == 0) {
            mbar_expect_tx(mbar + 1, (uint32_t)nq1 * 32u);
            if (nq1 > 0) bulk_g2s_hint(ts_s, A.quat + 4 * e0, (uint32_t)nq1 * 32u, mbar + 1, l2_policy_evict_first());
        }
        // ---- pass C: state recursion
        {
            double x0 = aex.a[0] * bc[13] + aex.b[0], x1 = aex.a[1] * bc[14] + aex.b[1], x2 = aex.a[2] * bc[15] + aex.b[2];
            if (c0 == 0) { zS[0] = bc[13]; zS[1] = bc[14]; zS[2] = bc[15]; }
#pragma unroll PASSC_UNROLL
            for (int i = s0; i < c1; ++i) {
                x0 = posS[3 * i] * x0 + zS[3 * i]; x1 = posS[3 * i + 1] * x1 + zS[3 * i + 1]; x2 = posS[3 * i + 2] * x2 + zS[3 * i + 2];
                zS[3 * i] = x0; zS[3 * i + 1] = x1; zS[3 * i + 2] = x2;
            }
        }
        fence_proxy_async();
        __syncwarp();
        // ---- store fused positions; stream the quaternions
        double* gout = A.out_pos + 3 * e0;
        if (viol_total) st |= ST_RANSAC_OUTLIERS;
        if (lane == 0) {
            const int m = n - lead, even = m & ~1;
            if (even > 0) { bulk_s2g_hint(gout + 3 * lead, zS + 3 * lead, (uint32_t)even * 24u, l2_policy_evict_first()); bulk_commit(); }
            if (lead) { gout[0] = zS[0]; gout[1] = zS[1]; gout[2] = zS[2]; }
            if (m & 1) {
                const int q = 3 * (lead + even);
                gout[q] = zS[q]; gout[q + 1] = zS[q + 1]; gout[q + 2] = zS[q + 2];
            }
            A.status[b] = st;
            if (A.sim3_out) {
                double* o = A.sim3_out + 16 * (size_t)b;
#pragma unroll
                for (int k = 0; k < 9; ++k) o[k] = bc[20 + k];
                o[9] = bc[16]; o[10] = bc[17]; o[11] = bc[18]; o[12] = bc[19];
                o[13] = (double)n; o[14] = (double)n; o[15] = (double)viol_total;
            }
            if (n > nq1) {
                mbar_expect_tx(mbar + 2, (uint32_t)(n - nq1) * 32u);
                bulk_g2s_hint(ts_s + 4 * nq1, A.quat + 4 * (e0 + nq1), (uint32_t)(n - nq1) * 32u, mbar + 2, l2_policy_evict_first());
            }
        }
        int bad = 0;
        {
            const Quat C{bc[9], bc[10], bc[11], bc[12]};
            const uint64_t pf = l2_policy_evict_first();
            double2* __restrict__ qout = reinterpret_cast<double2*>(A.out_quat + 4 * e0);
            const double2* __restrict__ q2 = reinterpret_cast<const double2*>(ts_s);
            mbar_wait_polite(mbar + 1, par_q0); par_q0 ^= 1;
            int i0 = 0, stop = min(n, nq1);
#pragma unroll 1
            for (int phase = 0; phase < 2; ++phase) {
#pragma unroll 1
                for (; i0 < stop; i0 += 32) {
                    const int i = i0 + lane;
                    const int ia = i < n ? i : 0;             // the tail re-reads a landed pose and stores nothing
                    const double2 lo0 = q2[2 * ia], hi0 = q2[2 * ia + 1];
                    const Quat qa{lo0.x, lo0.y, hi0.x, hi0.y};
                    const double na = qnorm2(qa);
                    if (na == 0.0) bad = 1;                   // scipy raises here (:466); output row becomes NaN
                    const Quat rr = qscale(qmul(C, qa), fast_rsqrt(na));
                    if (i < n) {
                        stg2_hint(qout + 2 * i, make_double2(rr.x, rr.y), pf);
                        stg2_hint(qout + 2 * i + 1, make_double2(rr.z, rr.w), pf);
                    }
                }
                if (phase == 0 && n > nq1) { mbar_wait_polite(mbar + 2, par_q1); par_q1 ^= 1; }
                stop = n;
            }
        }
        fence_proxy_async();                                  // generic reads of the buffers before the next TMA writes
        __syncwarp();                                         // (and lane 0's status store is ordered before the flag below)
        if (bad) atomicOr(A.status + b, ST_BAD_QUATERNION);
        if (lane == 0) bulk_wait_read();                      // the position store has left shared memory
        __syncwarp();
    }
}

template <int LCH>
static cudaError_t launch_warp_t(const FuseArgs& a, int num_sms, cudaStream_t stream) {
    const size_t smem = warp_smem_bytes(a.cap);
    auto kern = fuse_warp_kernel<LCH>;
    static std::atomic<long long> cached_smem{-1};
    static std::atomic<int> cached_per_sm{0};
    int per_sm = 0;
    if (cached_smem.load(std::memory_order_acquire) == (long long)smem) per_sm = cached_per_sm.load(std::memory_order_relaxed);
    else {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        e = cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        if (e != cudaSuccess) return e;
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, 32, smem);
        if (e != cudaSuccess) return e;
        if (per_sm < 1) return cudaErrorInvalidConfiguration;
        cached_per_sm.store(per_sm, std::memory_order_relaxed);
        cached_smem.store((long long)smem, std::memory_order_release);
    }
    long long grid = (long long)num_sms * per_sm;
    if (grid > a.B) grid = a.B;
    kern<<<(unsigned)grid, 32, smem, stream>>>(a);
    return cudaGetLastError();
}

// Deferred-trajectory counters (one per in-flight call, recycled round-robin): module-level device
// memory, so the *_dev entry point needs no workspace argument and allocates nothing.
__device__ int g_defer_count[256];          // per slot: deferred count, fast kernel work counter, general kernel chunk counter, pad
// One slot per in-flight call, handed out round-robin.  A slot is released by an event recorded behind the call's last
// kernel (defer_counter_release); a call that draws a slot whose previous user has not finished (more than 64 calls in
// flight on different streams, or overlapping replays of a captured graph) gets cudaErrorNotReady and the caller falls
// back to the general kernel, which needs no counters -- never a shared counter.
static cudaEvent_t g_defer_event[64];
static bool g_defer_event_ok[64];
static std::atomic<unsigned> g_defer_ticket{0};
cudaError_t defer_counter(int** out, int* slot_out) {
    int* base = nullptr;
    cudaError_t e = cudaGetSymbolAddress(reinterpret_cast<void**>(&base), g_defer_count);
    if (e != cudaSuccess) return e;
    const unsigned slot = g_defer_ticket.fetch_add(1) & 63u;
    if (g_defer_event_ok[slot]) {
        const cudaError_t q = cudaEventQuery(g_defer_event[slot]);
        if (q == cudaErrorNotReady) return cudaErrorNotReady;
        if (q != cudaSuccess) { cudaGetLastError(); }
    }
    *out = base + 4 * slot;
    *slot_out = (int)slot;
    return cudaSuccess;
}
void defer_counter_release(int slot, cudaStream_t stream) {
    cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing(stream, &cap) != cudaSuccess || cap != cudaStreamCaptureStatusNone) { cudaGetLastError(); return; }   // inside a graph capture: no host-side tracking
    if (!g_defer_event_ok[slot]) {
        if (cudaEventCreateWithFlags(&g_defer_event[slot], cudaEventDisableTiming) != cudaSuccess) { cudaGetLastError(); return; }
        g_defer_event_ok[slot] = true;
    }
    if (cudaEventRecord(g_defer_event[slot], stream) != cudaSuccess) cudaGetLastError();
}

template <int CT, int LCH>
static cudaError_t launch_fast_t(const FuseArgs& a, int num_sms, cudaStream_t stream) {
    size_t smem = fast_smem_bytes(a.cap, CT);
    if (const char* pad = getenv("GSF_FAST_PAD_SMEM")) smem += (size_t)atoi(pad);      // tuning hook: fewer blocks per SM
    auto kern = fuse_fast_kernel<CT, LCH>;
    // function attributes and occupancy are fixed per (instantiation, shared-memory size): looked up once, not per call
    static std::atomic<long long> cached_smem{-1};
    static std::atomic<int> cached_per_sm{0};
    int per_sm = 0;
    if (cached_smem.load(std::memory_order_acquire) == (long long)smem) per_sm = cached_per_sm.load(std::memory_order_relaxed);
    else {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        e = cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        if (e != cudaSuccess) return e;
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, CT + 64, smem);
        if (e != cudaSuccess) return e;
        if (per_sm < 1) return cudaErrorInvalidConfiguration;
        cached_per_sm.store(per_sm, std::memory_order_relaxed);
        cached_smem.store((long long)smem, std::memory_order_release);
    }
    long long grid = (long long)num_sms * per_sm;
    if (grid > a.B) grid = a.B;
    kern<<<(unsigned)grid, CT + 64, smem, stream>>>(a);
    return cudaGetLastError();
}

// Compute threads x chunk length per instantiation (chunk lengths odd: conflict-free 8-byte shared accesses).
// Variant id = compute threads * 100 + chunk length.
static int fast_variant(int cap) {
    const char* force = getenv("GSF_FAST_CT");                     // tuning hook
    const int f = force ? atoi(force) : 0;
    if (f == 128 && cap <= 128 * 9) return 12809;
    if (f == 96 && cap <= 96 * 11) return 9611;
    if (f == 64 && cap <= 64 * 9) return 6409;
    if (f == 64 && cap <= 64 * 17) return 6417;
    if (f == 32 && cap <= 32 * 9) return 3209;
    if (f == 1 && cap <= 32 * 9) return 1;           // one warp per trajectory (fuse_warp_kernel): measured slower, see there
    if (cap <= 32 * 9) return 3209;
    if (cap <= 64 * 9) return 6409;
    if (cap <= 64 * 17) return 6417;                 // two compute warps x 17 poses per thread: 1 % faster than 96 x 11 at 1000 poses
    // longer trajectories: general kernel (128 x 9 needs three 192-thread blocks per SM = 112 registers per thread and
    // measured no faster than the general kernel; it stays reachable through GSF_FAST_CT=128 for experiments)
    return 0;
}
static int variant_ct(int v) { return v / 100; }
bool fast_fuse_supported(int cap, int max_smem) {
    const int v = fast_variant(cap);
    if (v == 1) return warp_smem_bytes(cap) <= (size_t)max_smem;
    return v > 0 && fast_smem_bytes(cap, variant_ct(v)) <= (size_t)max_smem;
}
cudaError_t launch_fuse_fast(const FuseArgs& a, int num_sms, cudaStream_t stream) {
    switch (fast_variant(a.cap)) {
        case 1: return launch_warp_t<9>(a, num_sms, stream);
        case 3209: return launch_fast_t<32, 9>(a, num_sms, stream);
        case 6409: return launch_fast_t<64, 9>(a, num_sms, stream);
        case 6417: return launch_fast_t<64, 17>(a, num_sms, stream);
        case 9611: return launch_fast_t<96, 11>(a, num_sms, stream);
        case 12809: return launch_fast_t<128, 9>(a, num_sms, stream);
    }
    return cudaErrorInvalidValue;
}

}  // namespace gsf
