// gsf_fuse_batched: Sim3 point selection -> Umeyama -> EKF (+RTS) for a batch of independent
// trajectories, one thread block per trajectory at a time, the whole trajectory resident in
// shared memory so that HBM is read once (88 B/pose) and written once (56 B/pose).
//
// Reference path replaced (file:line in /root/reference/EKFGPSSLAM.py):
//   Sim3 point selection :972-998, compute_sim3_transform :428-459 (+ all-points residual
//   check standing in for the unseeded RANSAC :389-426), transform_trajectory row 0 :461-467,
//   apply_ekf_correction :831-935 (ExtendedKalmanFilter :679-772, RTS :777-803, sharp-turn
//   gate :808-826).
//
// Formulation (checked on the CPU by oracle/kernel_model.py against the step-by-step oracle):
//   * ts / positions / measurements of one trajectory are staged with three TMA bulk copies
//     (cp.async.bulk + mbarrier); the next trajectory's copies are issued as soon as the
//     fused positions of the current one have left shared memory, and overlap the streaming
//     quaternion pass; the trajectory after that is prefetched into L2.
//   * each thread owns a contiguous chunk of LCH poses (LCH odd => conflict-free 8-byte
//     shared-memory accesses across a warp);
//   * pass 1  validity flags + pivot-shifted Umeyama sums in one sweep (fast path: every pose
//     has GNSS, no time gap, inside the Sim3 window; otherwise the general selection path);
//     fixed-order chunk sums, xor-shuffle tree, warp-order combine => bit-reproducible;
//   * pass 2  covariance recursion p -> R(p+q)/(p+q+R) as Moebius maps, composed as rescaled
//     2x2 matrices (warp shuffle scan + warp-order combine), while warp 0 runs the 3x3
//     one-sided-Jacobi SVD;
//   * pass 3  exact per-step Joseph-form gains inside a chunk + affine state maps
//     x -> (1-k)(x+u) + k z (scan), residual check against the Sim3 fit;
//   * pass 4  state recursion, closed-form RTS over recovered outages
//     x_s[k] = x_f[k] + P_f[k]/P_pred[i] * delta_i;
//   * odometry is telescoped: M(q_state[i-1]) M(q_hat[i-1])^T == M(C), C = q_state0 (x)
//     conj(q_hat0), so u_i = M(C)(p_i - p_{i-1}) and q_state[i] = C (x) q_hat[i].
#include "gsf_fuse_shared.cuh"

namespace gsf {

// 1: quaternion rounds beyond the first run in the shadow of the next trajectory's SVD (measured slower: the
// next load is exposed and the SVD warp competes with them); 0: all rounds right after the store.
#ifndef GSF_DEFER_QUAT
#define GSF_DEFER_QUAT 0
#endif
#define GSF_STAMP(k) do { if (A.phase_clock && blockIdx.x == 0 && tid == 0 && it == 2) A.phase_clock[k] = clock64(); } while (0)

template <int ND> struct Pack { double v[ND]; };
// Exclusive block scan used by the (rare) general selection path.
template <int ND, class Op>
__device__ inline Pack<ND> block_exclusive_scan(Pack<ND> x, Op op, const Pack<ND>& ident, double* scratch) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        Pack<ND> y;
#pragma unroll
        for (int k = 0; k < ND; ++k) y.v[k] = __shfl_up_sync(GSF_FULL_MASK, x.v[k], o);
        if (lane >= o) x = op(y, x);
    }
    Pack<ND> excl;
#pragma unroll
    for (int k = 0; k < ND; ++k) excl.v[k] = __shfl_up_sync(GSF_FULL_MASK, x.v[k], 1);
    if (lane == 0) excl = ident;
    if (nwarp > 1) {
        __syncthreads();
        if (lane == 31) {
#pragma unroll
            for (int k = 0; k < ND; ++k) scratch[warp * ND + k] = x.v[k];
        }
        __syncthreads();
        Pack<ND> pre = ident;
        for (int w = 0; w < warp; ++w) {
            Pack<ND> tot;
#pragma unroll
            for (int k = 0; k < ND; ++k) tot.v[k] = scratch[w * ND + k];
            pre = op(pre, tot);
        }
        excl = op(pre, excl);
        __syncthreads();
    }
    return excl;
}
struct RankOp {     // v = [count, last valid timestamp (NaN = none)]
    __device__ Pack<2> operator()(const Pack<2>& e, const Pack<2>& l) const {
        Pack<2> r; r.v[0] = e.v[0] + l.v[0]; r.v[1] = isnan(l.v[1]) ? e.v[1] : l.v[1]; return r;
    }
};
__device__ __forceinline__ int block_min_int(int v, int* scratch) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = min(v, __shfl_xor_sync(GSF_FULL_MASK, v, o));
    if (nwarp == 1) return v;
    __syncthreads();
    if (lane == 0) scratch[warp] = v;
    __syncthreads();
    int r = scratch[0];
    for (int w = 1; w < nwarp; ++w) r = min(r, scratch[w]);
    __syncthreads();
    return r;
}

// Next trajectory of a block.  Whole batch: static stride.  Deferred pass (only the trajectories the fast kernel left):
// chunks of 8 consecutive indices are claimed from a global counter and the deferred ones among them are taken in order --
// a static stride gives every block a binomial share of the deferred trajectories (+-17 % at 50 %), and a load issued for
// the plain successor is wasted whenever that one is not deferred.  Called by thread 0 only; the cursor lives in its registers.
constexpr int DQ_CHUNK = 8;
struct TrajCursor { int base; unsigned mask; };
__device__ __forceinline__ int next_trajectory(const FuseArgs& A, int b, int stride, TrajCursor& cur) {
    if (!A.only_deferred || !A.work_counter) {
        int nb = b + stride;
        if (A.only_deferred) while (nb < A.B && A.status[nb] != ST_DEFERRED) nb += stride;
        return min(nb, A.B);
    }
    for (;;) {
        if (cur.mask) { const int k = __ffs(cur.mask) - 1; cur.mask &= cur.mask - 1; return cur.base + k; }
        const int c = atomicAdd(A.work_counter, DQ_CHUNK);
        if (c >= A.B) return A.B;
        int st[DQ_CHUNK];
#pragma unroll
        for (int k = 0; k < DQ_CHUNK; ++k) st[k] = c + k < A.B ? A.status[c + k] : 0;
        unsigned m = 0;
#pragma unroll
        for (int k = 0; k < DQ_CHUNK; ++k) if (st[k] == ST_DEFERRED) m |= 1u << k;
        cur.base = c; cur.mask = m;
    }
}

// ----------------------------------------------------------------------------- shared-memory map
constexpr int SM_SUMS = 0;          // 8 warps x 18 partial sums          (144)
constexpr int SM_MOEB = 144;        // 8 warps x 12 Moebius warp totals   (96)
constexpr int SM_AFF = 240;         // 8 warps x 6 affine warp totals     (48)
constexpr int SM_BC = 288;          // M(C) 0-8, C 9-12, x0 13-15, t 16-18, s 19, R 20-28 (32)
constexpr int SM_PRM = 320;         // FuseParams as 23 doubles           (24)
constexpr int SM_GEN = 344;         // general path: n, mu_s, mu_d, H, ss, t_first (24); its scans reuse SM_SUMS
constexpr int SM_DOUBLES = 400;
// ints: 0-7 block_min scratch, 8 status bits, 9 has-recovery, 10 residual violators,
//       11 general-path flag, 12 selection count, 13 valid count, 14 has-outage-step (gate), 15 next trajectory of the block

// Resident blocks per SM the register budget is sized for.
constexpr int fuse_min_blocks(int threads) { return threads <= 32 ? 14 : (threads == 64 ? 7 : (threads <= 160 ? 3 : 1)); }

template <int THREADS, int LCH>
__global__ void __launch_bounds__(THREADS, fuse_min_blocks(THREADS)) fuse_traj_kernel(const FuseArgs A) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    constexpr int NW = THREADS / 32;
    const int cap2 = (A.cap + 3) & ~1;                      // even => every sub-buffer stays 16-byte aligned
    double* ts_s = reinterpret_cast<double*>(smem_raw);
    double* pos_s = ts_s + cap2;
    double* z_s = pos_s + 3 * (size_t)cap2;
    double* sd = z_s + 3 * (size_t)cap2;                    // SM_DOUBLES scratch doubles
    int* iscr = reinterpret_cast<int*>(sd + SM_DOUBLES);    // 16 ints
    uint64_t* mbar = reinterpret_cast<uint64_t*>(iscr + 16);
    unsigned char* flg = reinterpret_cast<unsigned char*>(mbar + 2);
    double* bc = sd + SM_BC;
    const double* prm_s = sd + SM_PRM;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    uint32_t parity = 0;
    if (A.only_deferred && *A.defer_count == 0) return;     // nothing was left by the fast kernel
    const int stride = (int)gridDim.x;
    TrajCursor cursor{0, 0u};
    if (tid == 0) {
        mbar_init(mbar, 1); fence_mbar_init();
        const int f = next_trajectory(A, (int)blockIdx.x - stride, stride, cursor);      // (deferred pass: the block's first deferred trajectory)
        iscr[15] = f;
        if (A.use_tma && f < A.B) issue_trajectory_load(A, f, ts_s, pos_s, z_s, mbar);
    }
    __syncthreads();
    const int first = iscr[15];

    // quaternion rounds of the previous trajectory still to do (they run in the shadow of the next SVD)
    long long dq_e0 = 0; int dq_n = 0, dq_b = -1; Quat dq_C{0.0, 0.0, 0.0, 1.0};
    constexpr int DQ_T = NW > 1 ? THREADS - 32 : THREADS;      // threads that take part in the deferred rounds

    int it = 0;
    int nb = A.B;
    for (int b = first; b < A.B; b = nb, ++it) {
        const long long e0 = A.offsets[b];
        const int n = (int)(A.offsets[b + 1] - e0);
        const bool not_mine = A.only_deferred && A.status[b] != ST_DEFERRED;   // status[b] is only ever written by this block
        if (n <= 0 || n > A.cap || not_mine) {
            __syncthreads();                                // (everybody has read the previous iscr[15])
            if (tid == 0) {
                if (!not_mine) A.status[b] = n <= 0 ? ST_EMPTY : ST_TOO_LONG;
                const int nx = next_trajectory(A, b, stride, cursor);
                iscr[15] = nx;
                if (A.use_tma && nx < A.B) issue_trajectory_load(A, nx, ts_s, pos_s, z_s, mbar);
            }
            __syncthreads();
            nb = iscr[15];
            continue;
        }
        GSF_STAMP(0);
        // ------------------------------------------------------------------ stage inputs
        const int lead = A.use_tma ? (int)(e0 & 1) : 0;
        double* tsS = ts_s + lead; double* posS = pos_s + 3 * lead; double* zS = z_s + 3 * lead;
        if (tid < 23) sd[SM_PRM + tid] = reinterpret_cast<const double*>(A.params + (A.params_per_traj ? b : 0))[tid];
        if (tid == 32 % THREADS) { iscr[9] = 0; iscr[10] = 0; iscr[11] = 0; iscr[14] = 0; }
        if (A.use_tma) {
            const int cnt = n + lead, even = cnt & ~1;
            if ((cnt & 1) && tid < 7) {                       // odd tail element: plain copy
                const long long g = e0 - lead + even;
                if (tid == 0) ts_s[even] = A.ts[g];
                else if (tid < 4) pos_s[3 * even + (tid - 1)] = A.pos[3 * g + (tid - 1)];
                else z_s[3 * even + (tid - 4)] = A.z[3 * g + (tid - 4)];
            }
            if (even > 0) { mbar_wait(mbar, parity); parity ^= 1; }
        } else {
            for (int i = tid; i < n; i += THREADS) tsS[i] = A.ts[e0 + i];
            for (int i = tid; i < 3 * n; i += THREADS) { posS[i] = A.pos[3 * e0 + i]; zS[i] = A.z[3 * e0 + i]; }
        }
        __syncthreads();
        GSF_STAMP(1);

        const FuseParams& prm = *reinterpret_cast<const FuseParams*>(prm_s);
        const bool ekf_only = A.init_pos != nullptr;
        const int c0 = min(tid * LCH, n), c1 = min(c0 + LCH, n);
        const int s0 = max(c0, 1);                          // steps owned: i in [s0, c1)

        // ------------------------------------------------------------------ pass 1: flags + pivot-shifted Umeyama sums
        double pprev0 = 0.0, pprev1 = 0.0, pprev2 = 0.0;    // position of pose s0-1 (read before any in-place write)
        if (c0 < n) { const int ip = max(c0 - 1, 0); pprev0 = posS[3 * ip]; pprev1 = posS[3 * ip + 1]; pprev2 = posS[3 * ip + 2]; }
        {
            double v[17];
#pragma unroll
            for (int k = 0; k < 17; ++k) v[k] = 0.0;
            const double ps0 = posS[0], ps1 = posS[1], ps2 = posS[2];
            const double pz0 = zS[0], pz1 = zS[1], pz2 = zS[2];
            const double t_first = tsS[0], t_lim = t_first + prm.max_duration, gap = prm.gap_threshold;
            int viol = row_has_nan(pz0, pz1, pz2) ? 1 : 0, inval = 0;
            double tp = (c0 > 0 && c0 < n) ? tsS[c0 - 1] : t_first;
#pragma unroll 1
            for (int j = 0; j < LCH; ++j) {
                const int i = c0 + j;
                if (i < c1) {
                    const double z0 = zS[3 * i], z1 = zS[3 * i + 1], z2 = zS[3 * i + 2];
                    const double t = tsS[i];
                    const bool valid = !row_has_nan(z0, z1, z2);
                    flg[i] = valid ? FLAG_VALID : 0;
                    if (!valid) inval = 1;
                    if (!valid || t - tp > gap || t > t_lim) viol = 1;
                    tp = t;
                    if (valid && !ekf_only) {
                        const double a0 = posS[3 * i] - ps0, a1 = posS[3 * i + 1] - ps1, a2 = posS[3 * i + 2] - ps2;
                        const double b0 = z0 - pz0, b1 = z1 - pz1, b2 = z2 - pz2;
                        v[0] += 1.0; v[1] += a0; v[2] += a1; v[3] += a2; v[4] += b0; v[5] += b1; v[6] += b2;
                        v[7] += a0 * b0; v[8] += a0 * b1; v[9] += a0 * b2;
                        v[10] += a1 * b0; v[11] += a1 * b1; v[12] += a1 * b2;
                        v[13] += a2 * b0; v[14] += a2 * b1; v[15] += a2 * b2;
                        v[16] += a0 * a0 + a1 * a1 + a2 * a2;
                    }
                }
            }
            if (!ekf_only) {
                // xor-shuffle tree with the stage loop rolled (code size), the 17 sums unrolled
#pragma unroll 1
                for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
                    for (int k = 0; k < 17; ++k) v[k] += __shfl_xor_sync(GSF_FULL_MASK, v[k], o);
                }
                if (lane == 0) {
#pragma unroll
                    for (int k = 0; k < 17; ++k) sd[SM_SUMS + warp * 18 + k] = v[k];
                }
                if (viol) iscr[11] = 1;
            }
            if (inval) iscr[14] = 1;
        }
        __syncthreads();
        GSF_STAMP(2);
        int st = ST_OK;

        // ------------------------------------------------------------------ GNSS recoveries (:879-894): sharp-turn gate
        // is_sharp_turn_in_segment (:808-826) over the outage [s .. i-1] is a maximum of per-step yaw rates, and every step
        // costs nine transcendental calls (two as_euler('zyx') yaws, the wrap): the steps of all outages of the trajectory are
        // spread over the whole block (one thread walking a 60-pose outage alone held the block for ~100k cycles), each
        // leaves one flag bit, and the thread that owns the recovery pose ORs the bits of its outage.
        if (iscr[14]) {
            const double thr = prm.yaw_rate_thresh;
            const double* __restrict__ gq = A.quat + 4 * e0;
            for (int a = 1 + tid; a < n; a += THREADS) {
                if (!(flg[a] & FLAG_VALID) && !(flg[a - 1] & FLAG_VALID)) {
                    const double t1 = tsS[a - 1], t2 = tsS[a];
                    if (t2 <= t1) continue;
                    const double2 l1 = __ldg(reinterpret_cast<const double2*>(gq + 4 * (a - 1))), h1 = __ldg(reinterpret_cast<const double2*>(gq + 4 * (a - 1)) + 1);
                    const double2 l2 = __ldg(reinterpret_cast<const double2*>(gq + 4 * a)), h2 = __ldg(reinterpret_cast<const double2*>(gq + 4 * a) + 1);
                    const Quat q1{l1.x, l1.y, h1.x, h1.y}, q2{l2.x, l2.y, h2.x, h2.y};
                    bool sharp;
                    if (qnorm2(q1) == 0.0 || qnorm2(q2) == 0.0) sharp = true;          // scipy ValueError -> True (:821)
                    else {
                        const double y1 = yaw_zyx(q1), y2 = yaw_zyx(q2);
                        const double d = atan2(sin(y2 - y1), cos(y2 - y1));
                        sharp = fabs(d / (t2 - t1)) > thr;
                    }
                    if (sharp) flg[a] = (unsigned char)(flg[a] | FLAG_SHARP_STEP);     // bit 0 is unchanged: concurrent neighbour reads stay valid
                }
            }
            __syncthreads();
            GSF_STAMP(10);
            for (int i = s0; i < c1; ++i) {
                const int f = flg[i];
                if ((f & FLAG_VALID) && !(flg[i - 1] & FLAG_VALID)) {
                    int s = i - 1, sharp = 0;
                    while (s > 0 && !(flg[s - 1] & FLAG_VALID)) { sharp |= flg[s] & FLAG_SHARP_STEP; --s; }
                    int nf = f | FLAG_RECOVERY;
                    if (sharp) nf |= FLAG_NO_RTS;
                    flg[i] = (unsigned char)nf;             // bit 0 is unchanged: concurrent neighbour reads stay valid
                    iscr[9] = 1;
                }
            }
        }
        // ------------------------------------------------------------------ general Sim3 point selection (:972-998), rare
        GSF_STAMP(11);
        if (!ekf_only && iscr[11]) {
            int cntv = 0; double lastT = nan("");
            for (int i = c0; i < c1; ++i) if (flg[i] & FLAG_VALID) { ++cntv; lastT = tsS[i]; }
            Pack<2> mine; mine.v[0] = (double)cntv; mine.v[1] = lastT;
            Pack<2> id2; id2.v[0] = 0.0; id2.v[1] = nan("");
            Pack<2> pre = block_exclusive_scan<2>(mine, RankOp(), id2, sd + SM_SUMS);
            double tot[1] = {(double)cntv};
            block_sum<1>(tot, sd + SM_SUMS);
            const int nvalid = (int)tot[0];
            const int rank0 = (int)pre.v[0];
            int kmin = 0x7fffffff;
            {
                int r = rank0; double pt = pre.v[1];
                for (int i = c0; i < c1; ++i) if (flg[i] & FLAG_VALID) {
                    if (r == 0) sd[SM_GEN + 23] = tsS[i];               // timestamp of the first valid point
                    if (r >= 1 && tsS[i] - pt > prm.gap_threshold) kmin = min(kmin, r - 1);
                    pt = tsS[i]; ++r;
                }
            }
            kmin = block_min_int(kmin, iscr);
            __syncthreads();
            const int first_cnt = (kmin == 0x7fffffff) ? nvalid : kmin;
            int mode;                                   // 0 all valid, 1 first run, 2 first run within max_duration
            const double tlim = sd[SM_GEN + 23] + prm.max_duration;
            if (first_cnt < prm.min_samples) mode = 0;
            else {
                double timed[1] = {0.0};
                int r = rank0;
                for (int i = c0; i < c1; ++i) if (flg[i] & FLAG_VALID) { if (r < first_cnt && tsS[i] <= tlim) timed[0] += 1.0; ++r; }
                block_sum<1>(timed, sd + SM_SUMS);
                mode = ((int)timed[0] < prm.min_samples) ? 1 : 2;
            }
            double s7[7] = {0, 0, 0, 0, 0, 0, 0};
            {
                int r = rank0;
                for (int i = c0; i < c1; ++i) if (flg[i] & FLAG_VALID) {
                    const bool sel = mode == 0 || (r < first_cnt && (mode == 1 || tsS[i] <= tlim));
                    if (sel) {
                        flg[i] |= FLAG_SELECTED;
                        s7[0] += posS[3 * i]; s7[1] += posS[3 * i + 1]; s7[2] += posS[3 * i + 2];
                        s7[3] += zS[3 * i]; s7[4] += zS[3 * i + 1]; s7[5] += zS[3 * i + 2];
                        s7[6] += 1.0;
                    }
                    ++r;
                }
            }
            block_sum<7>(s7, sd + SM_SUMS);
            const double inv_n = 1.0 / fmax(s7[6], 1.0);
            const double m0 = s7[0] * inv_n, m1 = s7[1] * inv_n, m2 = s7[2] * inv_n;
            const double d0 = s7[3] * inv_n, d1 = s7[4] * inv_n, d2 = s7[5] * inv_n;
            double h[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
            for (int i = c0; i < c1; ++i) if (flg[i] & FLAG_SELECTED) {
                const double a0 = posS[3 * i] - m0, a1 = posS[3 * i + 1] - m1, a2 = posS[3 * i + 2] - m2;
                const double b0 = zS[3 * i] - d0, b1 = zS[3 * i + 1] - d1, b2 = zS[3 * i + 2] - d2;
                h[0] += a0 * b0; h[1] += a0 * b1; h[2] += a0 * b2;
                h[3] += a1 * b0; h[4] += a1 * b1; h[5] += a1 * b2;
                h[6] += a2 * b0; h[7] += a2 * b1; h[8] += a2 * b2;
                h[9] += a0 * a0 + a1 * a1 + a2 * a2;
            }
            block_sum<10>(h, sd + SM_SUMS);
            __syncthreads();
            if (tid == 0) {
                double* g = sd + SM_GEN;
                g[0] = s7[6]; g[1] = m0; g[2] = m1; g[3] = m2; g[4] = d0; g[5] = d1; g[6] = d2;
#pragma unroll
                for (int k = 0; k < 10; ++k) g[7 + k] = h[k];
                iscr[13] = nvalid;
            }
            __syncthreads();
        }

        // ------------------------------------------------------------------ pass 2: SVD (warp 0) | Moebius maps + deferred quaternions
        const bool xy_same = prm.p0[0] == prm.p0[1] && prm.q[0] == prm.q[1] && prm.r[0] == prm.r[1];
        Moeb3 mex;                                          // exclusive prefix inside the warp
        GSF_STAMP(7);
        if (warp == 0) {
            int ust = 0;
            if (!ekf_only) {
                double n_sel, ms_[3], md_[3], hh[9], ss;
                if (iscr[11]) {                             // general path: centred sums are ready
                    const double* g = sd + SM_GEN;
                    n_sel = g[0];
#pragma unroll
                    for (int k = 0; k < 3; ++k) { ms_[k] = g[1 + k]; md_[k] = g[4 + k]; }
#pragma unroll
                    for (int k = 0; k < 9; ++k) hh[k] = g[7 + k];
                    ss = g[16];
                } else {                                    // fast path: combine warp partials in warp order, un-shift the pivot
                    double v[17];
#pragma unroll
                    for (int k = 0; k < 17; ++k) { v[k] = sd[SM_SUMS + k]; for (int w = 1; w < NW; ++w) v[k] += sd[SM_SUMS + w * 18 + k]; }
                    n_sel = v[0];
                    const double inv = 1.0 / fmax(n_sel, 1.0);
                    const double ma[3] = {v[1] * inv, v[2] * inv, v[3] * inv}, mb[3] = {v[4] * inv, v[5] * inv, v[6] * inv};
#pragma unroll
                    for (int k = 0; k < 3; ++k) { ms_[k] = posS[k] + ma[k]; md_[k] = zS[k] + mb[k]; }
#pragma unroll
                    for (int r = 0; r < 3; ++r)
#pragma unroll
                        for (int c = 0; c < 3; ++c) hh[3 * r + c] = v[7 + 3 * r + c] - n_sel * ma[r] * mb[c];
                    ss = v[16] - n_sel * (ma[0] * ma[0] + ma[1] * ma[1] + ma[2] * ma[2]);
                }
                const int nsel = (int)n_sel;
                double R[9], t[3], s = 1.0;
                GSF_STAMP(8);
                if (nsel < 3 || nsel < prm.min_samples) ust |= ST_TOO_FEW_POINTS;
                else ust |= umeyama_finish_ool(nsel, ms_, md_, hh, ss, R, t, &s);
                GSF_STAMP(9);
                const Quat q0{A.quat[4 * e0], A.quat[4 * e0 + 1], A.quat[4 * e0 + 2], A.quat[4 * e0 + 3]};
                if (qnorm2(q0) == 0.0) ust |= ST_BAD_QUATERNION;
                if (lane == 0) {
                    if (!(ust & ST_TOO_FEW_POINTS)) {
                        const Quat qR = quat_from_matrix(R);
                        const Quat q0h = qunit(q0);
                        const Quat qs0 = qunit_or_identity(qmul(qR, q0h));
                        const Quat Cq = qmul(qs0, qconj(q0h));
                        double M[9]; qmat(Cq, M);
#pragma unroll
                        for (int k = 0; k < 9; ++k) { bc[k] = M[k]; bc[20 + k] = R[k]; }
                        bc[9] = Cq.x; bc[10] = Cq.y; bc[11] = Cq.z; bc[12] = Cq.w;
                        double rx, ry, rz;
                        mat_vec(R, posS[0], posS[1], posS[2], rx, ry, rz);
                        bc[13] = s * rx + t[0]; bc[14] = s * ry + t[1]; bc[15] = s * rz + t[2];
                        bc[16] = t[0]; bc[17] = t[1]; bc[18] = t[2]; bc[19] = s;
                    }
                    iscr[8] = ust; iscr[12] = nsel;
                    if (!iscr[11]) iscr[13] = nsel;
                }
            } else if (lane == 0) {
                const Quat q0{A.quat[4 * e0], A.quat[4 * e0 + 1], A.quat[4 * e0 + 2], A.quat[4 * e0 + 3]};
                const Quat qi{A.init_quat[4 * b], A.init_quat[4 * b + 1], A.init_quat[4 * b + 2], A.init_quat[4 * b + 3]};
                ust = (qnorm2(q0) == 0.0) ? ST_BAD_QUATERNION : 0;
                const Quat qs0 = qunit_or_identity(qi);
                const Quat Cq = qmul(qs0, qconj(qunit(q0)));
                double M[9]; qmat(Cq, M);
#pragma unroll
                for (int k = 0; k < 9; ++k) bc[k] = M[k];
                bc[9] = Cq.x; bc[10] = Cq.y; bc[11] = Cq.z; bc[12] = Cq.w;
                bc[13] = A.init_pos[3 * b]; bc[14] = A.init_pos[3 * b + 1]; bc[15] = A.init_pos[3 * b + 2];
                iscr[8] = ust;
            }
        }
        mex = xy_same ? moebius_chunk_scan<2>(tsS, flg, prm, s0, c1, lane, sd + SM_MOEB + warp * 12)
                      : moebius_chunk_scan<3>(tsS, flg, prm, s0, c1, lane, sd + SM_MOEB + warp * 12);
        __syncthreads();
        GSF_STAMP(3);
        st |= iscr[8];
        if (st & (ST_TOO_FEW_POINTS | ST_BAD_QUATERNION)) {
            // The reference aborts the run here (ValueError / RuntimeError): outputs are NaN.
            for (int i = tid; i < 3 * n; i += THREADS) A.out_pos[3 * e0 + i] = nan("");
            for (int i = tid; i < 4 * n; i += THREADS) A.out_quat[4 * e0 + i] = nan("");
            if (tid == 0) {
                A.status[b] = st;
                if (A.sim3_out) {
                    double* o = A.sim3_out + 16 * (size_t)b;
                    for (int k = 0; k < 13; ++k) o[k] = nan("");
                    o[13] = (double)iscr[12]; o[14] = (double)iscr[13]; o[15] = 0.0;
                }
                const int nx = next_trajectory(A, b, stride, cursor);
                iscr[15] = nx;
                if (A.use_tma && nx < A.B) issue_trajectory_load(A, nx, ts_s, pos_s, z_s, mbar);
            }
            __syncthreads();
            nb = iscr[15];
            continue;
        }

        // ------------------------------------------------------------------ pass 3: gains, affine maps, residual check
        Aff3 aex;                                           // exclusive affine prefix inside the warp
        {
            Moeb3 pre = mex;
            if (NW > 1 && warp > 0) {
                Moeb3 acc;
#pragma unroll
                for (int k = 0; k < 12; ++k) acc.m[k] = sd[SM_MOEB + k];
                for (int w = 1; w < warp; ++w) {
                    Moeb3 nx;
#pragma unroll
                    for (int k = 0; k < 12; ++k) nx.m[k] = sd[SM_MOEB + w * 12 + k];
                    acc = moeb_compose<3>(acc, nx);
                }
                pre = moeb_compose<3>(acc, mex);
            }
            double P[3];
#pragma unroll
            for (int a = 0; a < 3; ++a) {
                const double* m = pre.m + 4 * a;
                P[a] = (m[0] * prm.p0[a] + m[1]) / (m[2] * prm.p0[a] + m[3]);
            }
        // (pose 0's own position was consumed in pass 1 as pivot and is re-read from global below)
            double RC[9];
#pragma unroll
            for (int k = 0; k < 9; ++k) RC[k] = bc[k];
            const double sc = bc[19], t0 = bc[16], t1 = bc[17], t2 = bc[18];
            const double thr2 = (ekf_only || !(prm.residual_thresh > 0.0)) ? -1.0 : prm.residual_thresh * prm.residual_thresh;
            const bool general = iscr[11] != 0;
            double w_sharp = 1.0;
            if (prm.sharp_turn_steps > 0) { const double wd = 1.0 / (double)prm.sharp_turn_steps; if (wd < 1.0) w_sharp = wd; }
            int nviol = 0;
            Aff3 aff;
#pragma unroll
            for (int a = 0; a < 3; ++a) { aff.a[a] = 1.0; aff.b[a] = 0.0; }
            if (c0 == 0 && c1 > 0) {
                if (thr2 > 0.0 && (flg[0] & FLAG_VALID) && (!general || (flg[0] & FLAG_SELECTED))) {
                    double rx, ry, rz;
                    mat_vec(RC, posS[0], posS[1], posS[2], rx, ry, rz);
                    const double d0 = sc * rx + t0 - zS[0], d1 = sc * ry + t1 - zS[1], d2 = sc * rz + t2 - zS[2];
                    if (!(d0 * d0 + d1 * d1 + d2 * d2 < thr2)) ++nviol;
                }
            }
            // y = s * M(C) * p(s0-1) + t, advanced by s*u each step: the Sim3 image used by the residual check
            double y0 = 0.0, y1 = 0.0, y2 = 0.0;
            if (thr2 > 0.0 && s0 < c1) {
                mat_vec(RC, pprev0, pprev1, pprev2, y0, y1, y2);
                y0 = sc * y0 + t0; y1 = sc * y1 + t1; y2 = sc * y2 + t2;
            }
            const double q0 = prm.q[0], q1 = prm.q[1], q2 = prm.q[2], r0 = prm.r[0], r1 = prm.r[1], r2 = prm.r[2];
#pragma unroll 1
            for (int i = s0; i < c1; ++i) {
                const double dt = fmax(1e-6, tsS[i] - tsS[i - 1]);
                const double p0 = posS[3 * i], p1 = posS[3 * i + 1], p2 = posS[3 * i + 2];
                double u[3];
                mat_vec(RC, p0 - pprev0, p1 - pprev1, p2 - pprev2, u[0], u[1], u[2]);
                pprev0 = p0; pprev1 = p1; pprev2 = p2;
                const int f = flg[i];
                const double qq[3] = {q0 * dt, q1 * dt, q2 * dt};
                if (f & FLAG_VALID) {
                    const double zz[3] = {zS[3 * i], zS[3 * i + 1], zS[3 * i + 2]};
                    if (thr2 > 0.0) {
                        y0 = fma(sc, u[0], y0); y1 = fma(sc, u[1], y1); y2 = fma(sc, u[2], y2);
                        const double d0 = y0 - zz[0], d1 = y1 - zz[1], d2 = y2 - zz[2];
                        if ((!general || (f & FLAG_SELECTED)) && !(d0 * d0 + d1 * d1 + d2 * d2 < thr2)) ++nviol;
                    }
                    const double rr[3] = {r0, r1, r2};
                    double kk[3], om[3];
#pragma unroll
                    for (int a = 0; a < 3; ++a) {
                        if (a == 1 && xy_same) { kk[1] = kk[0]; om[1] = om[0]; P[1] = P[0]; continue; }
                        const double pp = P[a] + qq[a];
                        kk[a] = pp * fast_rcp(pp + rr[a]);
                        om[a] = 1.0 - kk[a];
                        P[a] = om[a] * pp * om[a] + kk[a] * rr[a] * kk[a];      // Joseph form (:731)
                    }
                    if (f & FLAG_NO_RTS) {                   // blended update at a sharp-turn recovery (:754-767)
#pragma unroll
                        for (int a = 0; a < 3; ++a) { kk[a] *= w_sharp; om[a] = 1.0 - kk[a]; }
                    }
#pragma unroll
                    for (int a = 0; a < 3; ++a) {
                        const double bv = om[a] * u[a] + kk[a] * zz[a];
                        posS[3 * i + a] = om[a]; zS[3 * i + a] = bv;
                        aff.b[a] = om[a] * aff.b[a] + bv; aff.a[a] *= om[a];
                    }
                } else {
                    if (thr2 > 0.0) { y0 = fma(sc, u[0], y0); y1 = fma(sc, u[1], y1); y2 = fma(sc, u[2], y2); }
#pragma unroll
                    for (int a = 0; a < 3; ++a) {
                        P[a] += qq[a];
                        posS[3 * i + a] = P[a];                          // P_f[i] kept for the RTS patch
                        zS[3 * i + a] = u[a];
                        aff.b[a] += u[a];
                    }
                }
            }
            if (c0 == 0 && c1 > 0) { posS[0] = prm.p0[0]; posS[1] = prm.p0[1]; posS[2] = prm.p0[2]; }   // P_f[0] for the RTS patch
            if (thr2 > 0.0) {
                nviol = warp_sum_i(nviol);
                if (lane == 0 && nviol) atomicAdd(iscr + 10, nviol);
            }
            aff_warp_scan(aff, lane);
            if (lane == 31) {
#pragma unroll
                for (int k = 0; k < 3; ++k) { sd[SM_AFF + warp * 6 + k] = aff.a[k]; sd[SM_AFF + warp * 6 + 3 + k] = aff.b[k]; }
            }
#pragma unroll
            for (int k = 0; k < 3; ++k) { aex.a[k] = __shfl_up_sync(GSF_FULL_MASK, aff.a[k], 1); aex.b[k] = __shfl_up_sync(GSF_FULL_MASK, aff.b[k], 1); }
            if (lane == 0) {
#pragma unroll
                for (int k = 0; k < 3; ++k) { aex.a[k] = 1.0; aex.b[k] = 0.0; }
            }
        }
        __syncthreads();
        GSF_STAMP(4);

        // ------------------------------------------------------------------ pass 4: state recursion
        {
            Aff3 pre = aex;
            if (NW > 1 && warp > 0) {
                Aff3 acc;
#pragma unroll
                for (int k = 0; k < 3; ++k) { acc.a[k] = sd[SM_AFF + k]; acc.b[k] = sd[SM_AFF + 3 + k]; }
                for (int w = 1; w < warp; ++w) {
                    Aff3 nx;
#pragma unroll
                    for (int k = 0; k < 3; ++k) { nx.a[k] = sd[SM_AFF + w * 6 + k]; nx.b[k] = sd[SM_AFF + w * 6 + 3 + k]; }
                    acc = aff_compose(acc, nx);
                }
                pre = aff_compose(acc, aex);
            }
            double x0 = pre.a[0] * bc[13] + pre.b[0], x1 = pre.a[1] * bc[14] + pre.b[1], x2 = pre.a[2] * bc[15] + pre.b[2];
            if (c0 == 0 && c1 > 0) { zS[0] = bc[13]; zS[1] = bc[14]; zS[2] = bc[15]; }
            for (int i = s0; i < c1; ++i) {
                if (flg[i] & FLAG_VALID) {
                    x0 = posS[3 * i] * x0 + zS[3 * i]; x1 = posS[3 * i + 1] * x1 + zS[3 * i + 1]; x2 = posS[3 * i + 2] * x2 + zS[3 * i + 2];
                } else {
                    x0 += zS[3 * i]; x1 += zS[3 * i + 1]; x2 += zS[3 * i + 2];
                }
                zS[3 * i] = x0; zS[3 * i + 1] = x1; zS[3 * i + 2] = x2;
            }
        }
        // ------------------------------------------------------------------ closed-form RTS over recovered outages
        // x_s[k] = x_f[k] + P_f[k] / P_pred[i] (x_f[i] - x_pred[i]) for the poses k of an outage recovered at i (:785-799 collapsed,
        // see the header).  Two block-wide steps instead of the recovery thread patching its whole outage alone: (1) the owner
        // of a recovery pose i leaves g = (x_f[i] - x_pred[i]) / P_pred[i] in the covariance row of pose i (free after pass C:
        // RTS reads the covariance rows of outage poses only); (2) every thread patches the outage poses of its own chunk,
        // walking it backwards with the g of the recovery that follows (found by a short look-ahead for the chunk's last run).
        if (iscr[9]) {
            __syncthreads();
            for (int i = s0; i < c1; ++i) {
                const int f = flg[i];
                if ((f & FLAG_RECOVERY) && !(f & FLAG_NO_RTS)) {
                    const double dt = fmax(1e-6, tsS[i] - tsS[i - 1]);
                    const double* gp = A.pos + 3 * (e0 + i);
                    double u[3];
                    mat_vec(bc, gp[0] - gp[-3], gp[1] - gp[-2], gp[2] - gp[-1], u[0], u[1], u[2]);
#pragma unroll
                    for (int a = 0; a < 3; ++a) {
                        const double ratio_den = posS[3 * (i - 1) + a] + prm.q[a] * dt;          // P_pred[i]
                        const double delta = zS[3 * i + a] - (zS[3 * (i - 1) + a] + u[a]);        // x_f[i] - x_pred[i]
                        posS[3 * i + a] = delta / ratio_den;
                    }
                }
            }
            __syncthreads();
            if (c0 < c1) {
                bool active = false;
                double g0 = 0.0, g1 = 0.0, g2 = 0.0;
                if (!(flg[c1 - 1] & FLAG_VALID)) {                     // the chunk ends inside an outage: its recovery lies ahead
                    int j = c1;
                    while (j < n && !(flg[j] & FLAG_VALID)) ++j;
                    if (j < n && (flg[j] & FLAG_RECOVERY) && !(flg[j] & FLAG_NO_RTS)) { active = true; g0 = posS[3 * j]; g1 = posS[3 * j + 1]; g2 = posS[3 * j + 2]; }
                }
                for (int k = c1 - 1; k >= c0; --k) {
                    const int f = flg[k];
                    if (f & FLAG_VALID) {
                        active = (f & FLAG_RECOVERY) && !(f & FLAG_NO_RTS);
                        if (active) { g0 = posS[3 * k]; g1 = posS[3 * k + 1]; g2 = posS[3 * k + 2]; }
                    } else if (active) {
                        zS[3 * k] += posS[3 * k] * g0; zS[3 * k + 1] += posS[3 * k + 1] * g1; zS[3 * k + 2] += posS[3 * k + 2] * g2;
                    }
                }
            }
        }
        if (A.use_tma) fence_proxy_async();
        __syncthreads();
        GSF_STAMP(5);

        // ------------------------------------------------------------------ store fused positions; stream the quaternions
        // q_state[i] = C (x) q_hat[i], global -> global in rounds of 4 poses per thread.  Thread 0
        // issues the bulk store first, and after the first round (when the store has drained out of
        // shared memory) the next trajectory's bulk loads, which then overlap the remaining rounds.
        double* gout = A.out_pos + 3 * e0;
        if (iscr[10]) st |= ST_RANSAC_OUTLIERS;
        if (A.use_tma) {
            if (tid == 0) {
                const int m = n - lead, even = m & ~1;
                if (even > 0) { bulk_s2g(gout + 3 * lead, zS + 3 * lead, (uint32_t)even * 24u); bulk_commit(); }
                if (lead) { gout[0] = zS[0]; gout[1] = zS[1]; gout[2] = zS[2]; }
                if (m & 1) {
                    const int q = 3 * (lead + even);
                    gout[q] = zS[q]; gout[q + 1] = zS[q + 1]; gout[q + 2] = zS[q + 2];
                }
            }
        } else {
            for (int i = tid; i < 3 * n; i += THREADS) gout[i] = zS[i];
        }
        if (tid == 0) {
            A.status[b] = st;
            if (A.sim3_out && !ekf_only) {
                double* o = A.sim3_out + 16 * (size_t)b;
#pragma unroll
                for (int k = 0; k < 9; ++k) o[k] = bc[20 + k];
                o[9] = bc[16]; o[10] = bc[17]; o[11] = bc[18]; o[12] = bc[19];
                o[13] = (double)iscr[12]; o[14] = (double)iscr[13]; o[15] = (double)iscr[10];
            }
        }
        {
            const Quat C{bc[9], bc[10], bc[11], bc[12]};
            // first round (4 poses per thread) now, while the bulk store drains ...
            const int n_now = (NW > 1 && GSF_DEFER_QUAT) ? min(n, 4 * THREADS) : n;
            int bad = quat_rounds(A.quat + 4 * e0, A.out_quat + 4 * e0, C, tid, THREADS, n_now);
            if (tid == 0) {
                const int nx = next_trajectory(A, b, stride, cursor);
                iscr[15] = nx;
                if (A.use_tma) {
                    bulk_wait_read();                               // shared memory is free again
                    fence_proxy_async();
                    if (nx < A.B) issue_trajectory_load(A, nx, ts_s, pos_s, z_s, mbar);
                }
            }
            // ... the remaining rounds run in the shadow of the next trajectory's SVD (or after the loop)
            dq_e0 = e0; dq_n = (NW > 1 && GSF_DEFER_QUAT) ? n : 0; dq_b = b; dq_C = C;
            __syncthreads();                                        // status[b] is written; scratch may be reused
            nb = iscr[15];
            if (bad) atomicOr(A.status + b, ST_BAD_QUATERNION);
        }
        GSF_STAMP(6);
    }
    if (dq_n > 4 * THREADS) {                                       // flush: last trajectory of this block
        const int bad = quat_rounds(A.quat + 4 * dq_e0, A.out_quat + 4 * dq_e0, dq_C, 4 * THREADS + tid, THREADS, dq_n);
        if (bad) atomicOr(A.status + dq_b, ST_BAD_QUATERNION);
    }
}

// ----------------------------------------------------------------------------- strict batched kernel
// One thread per trajectory running the literal step-by-step recursion (gsf_ekf_strict.cuh)
// straight from global memory.  Reference surface: apply_ekf_correction for arbitrary inputs
// (keeps the zero-motion fallback of :84-86); also the independent on-device cross-check of
// the scan formulation above.
__global__ void ekf_strict_kernel(const double* ts, const double* pos, const double* quat, const double* z,
                                  const long long* offsets, const FuseParams* params, int params_per_traj,
                                  const double* init_pos, const double* init_quat,
                                  double* out_pos, double* out_quat, int* status, int B) {
    int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    long long e0 = offsets[b];
    long n = (long)(offsets[b + 1] - e0);
    status[b] = ekf_strict_trajectory(n, ts + e0, pos + 3 * e0, quat + 4 * e0, z + 3 * e0,
                                      init_pos + 3 * b, init_quat + 4 * b, params[params_per_traj ? b : 0],
                                      out_pos + 3 * e0, out_quat + 4 * e0);
}

size_t fuse_smem_bytes(int cap) {
    const size_t cap2 = (size_t)((cap + 3) & ~1);
    return cap2 * 56 + SM_DOUBLES * 8 + 16 * 4 + 16 + ((cap2 + 15) & ~(size_t)15);
}

template <int THREADS, int LCH>
static cudaError_t launch_fuse_t(const FuseArgs& a, int num_sms, cudaStream_t stream) {
    const size_t smem = fuse_smem_bytes(a.cap);
    auto kern = fuse_traj_kernel<THREADS, LCH>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    if (e != cudaSuccess) return e;
    int per_sm = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, THREADS, smem);
    if (e != cudaSuccess) return e;
    if (per_sm < 1) return cudaErrorInvalidConfiguration;
    long long grid = (long long)num_sms * per_sm;
    if (grid > a.B) grid = a.B;
    kern<<<(unsigned)grid, THREADS, smem, stream>>>(a);
    return cudaGetLastError();
}

// threads x chunk must cover the longest trajectory; chunk lengths are odd.
cudaError_t launch_fuse(const FuseArgs& a, int threads, int num_sms, cudaStream_t stream) {
    const int need = (a.cap + threads - 1) / threads;
    if (threads == 160) return need <= 7 ? launch_fuse_t<160, 7>(a, num_sms, stream) : cudaErrorInvalidValue;
    if (need <= 9) {
        switch (threads) {
            case 32: return launch_fuse_t<32, 9>(a, num_sms, stream);
            case 64: return launch_fuse_t<64, 9>(a, num_sms, stream);
            case 128: return launch_fuse_t<128, 9>(a, num_sms, stream);
            case 256: return launch_fuse_t<256, 9>(a, num_sms, stream);
        }
    } else if (need <= 17 && threads == 256) {
        return launch_fuse_t<256, 17>(a, num_sms, stream);
    } else if (need <= 33 && threads == 256) {
        return launch_fuse_t<256, 33>(a, num_sms, stream);
    }
    return cudaErrorInvalidValue;
}

cudaError_t launch_ekf_strict(const double* ts, const double* pos, const double* quat, const double* z,
                              const long long* offsets, const FuseParams* params, int params_per_traj,
                              const double* init_pos, const double* init_quat, double* out_pos, double* out_quat,
                              int* status, int B, cudaStream_t stream) {
    if (B <= 0) return cudaSuccess;
    ekf_strict_kernel<<<(B + 63) / 64, 64, 0, stream>>>(ts, pos, quat, z, offsets, params, params_per_traj,
                                                        init_pos, init_quat, out_pos, out_quat, status, B);
    return cudaGetLastError();
}

}  // namespace gsf
