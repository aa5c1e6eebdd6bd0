// gsf_fuse_batched, long-trajectory kernel: the fused Sim3 -> EKF (+RTS) path of gsf_fused.cu for
// trajectories that do not fit the shared-memory staging of the general kernel (more than ~4000 poses;
// KITTI 00 / 02 / 08 are 4541 / 4661 / 4071).  The reference has no length limit
// (/root/reference/EKFGPSSLAM.py:831-935, :972-998, :428-459), so neither has this path.
//
// One block of 256 threads owns one trajectory at a time and streams it through shared memory in tiles
// of 2304 poses (9 per thread), carrying the covariance and the state from tile to tile:
//   phase A  validity flags, the general Sim3 point selection (:972-998, incl. its slice end at the
//            first gap) and the two-pass centred Umeyama sums, as pose-strided sweeps over global
//            memory (a trajectory of this size sits in L2 after the first sweep); fixed summation
//            order => bit-reproducible and independent of the batch sharding;
//   phase B  3x3 one-sided Jacobi SVD, R / t / s, C = q_state0 (x) conj(q_hat0) on warp 0;
//   phase C  per tile: Moebius covariance scan seeded with the carried covariance, exact per-step gains,
//            affine state scan seeded with the carried state (the same chunk formulation as the general
//            kernel), fused positions to out_pos;
//   phase D  closed-form RTS patches over recovered outages (:875-928; they may span tiles) and the
//            sharp-turn gate (:808-826), on out_pos in global memory;
//   phase E  quaternion pass q_state[i] = C (x) q_hat[i].
// Scratch: the kernel allocates nothing; until phase E, out_quat holds per pose the filtered covariance
// of outage poses (3 doubles) and the pose flags (1 byte).
#include "gsf_fuse_shared.cuh"

namespace gsf {

constexpr int LT = 256;                  // threads per block
constexpr int LNW = LT / 32;
constexpr int LLCH = 9;                  // poses per thread and tile (odd: conflict-free 8-byte accesses)
constexpr int LTILE = LT * LLCH;         // 2304 poses per tile

// shared-memory map (doubles after the staging buffers)
constexpr int LS_SUMS = 0;               // block_sum scratch, 8 warps x 10 (144 reserved)
constexpr int LS_MOEB = 144;             // 8 warps x 12 Moebius warp totals
constexpr int LS_AFF = 240;              // 8 warps x 6 affine warp totals
constexpr int LS_BC = 288;               // M(C) 0-8, C 9-12, x0 13-15, t 16-18, s 19, R 20-28
constexpr int LS_PRM = 320;              // FuseParams (23 doubles)
constexpr int LS_CARRY = 344;            // covariance (3) and state (3) at the end of the previous tile
constexpr int LS_GEN = 352;              // selection: n_sel, means (6), H (9), ss
constexpr int LS_WARP = 376;             // S1: per warp last / first valid timestamp (16)
constexpr int LS_DOUBLES = 400;
// ints: 0-7 warp counts, 8-15 warp last valid index, 16 carried count, 17 carried last index, 18 status,
//       19 has-recovery, 20 residual violators, 21 n_sel, 22 n_valid, 23 i_cut, 24 mode

__host__ __device__ constexpr size_t long_smem_bytes() {
    return (size_t)(LTILE + 2) * 8 + (size_t)3 * (LTILE + 2) * 8 + (size_t)3 * LTILE * 8 + (size_t)LS_DOUBLES * 8 + 32 * 4 + 16 +
           (size_t)(LTILE + 32);
}

__global__ void __launch_bounds__(LT, 1) fuse_long_kernel(const FuseArgs A) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double* const ts_s = reinterpret_cast<double*>(smem_raw);          // [LTILE + 2], element 0 = pose before the tile
    double* const pos_s = ts_s + (LTILE + 2);                          // [3 (LTILE + 2)], row 0 = pose before the tile
    double* const z_s = pos_s + 3 * (LTILE + 2);                       // [3 LTILE]
    double* const sd = z_s + 3 * LTILE;
    int* const iscr = reinterpret_cast<int*>(sd + LS_DOUBLES);         // 32 ints
    unsigned long long* const gap64 = reinterpret_cast<unsigned long long*>(iscr + 32);
    unsigned char* const flg_s = reinterpret_cast<unsigned char*>(gap64 + 2);      // [16 + LTILE], byte 15 = pose before the tile
    double* const bc = sd + LS_BC;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    for (int b = blockIdx.x; b < A.B; b += gridDim.x) {
        const long long e0 = A.offsets[b];
        const long long nn = A.offsets[b + 1] - e0;
        if (nn <= A.cap) continue;                          // handled by the shared-memory kernels
        const int n = (int)nn;
        const double* __restrict__ gts = A.ts + e0;
        const double* __restrict__ gpos = A.pos + 3 * e0;
        const double* __restrict__ gz = A.z + 3 * e0;
        double* const gout = A.out_pos + 3 * e0;
        double* const gscr = A.out_quat + 4 * e0;           // scratch until phase E: [4i .. 4i+2] = P_f[i], byte 0 of [4i+3] = flags
        auto gflag = [&](int i) -> unsigned char& { return reinterpret_cast<unsigned char*>(gscr + 4 * (size_t)i + 3)[0]; };

        __syncthreads();                                    // previous trajectory is done with the scratch
        if (tid < 23) sd[LS_PRM + tid] = reinterpret_cast<const double*>(A.params + (A.params_per_traj ? b : 0))[tid];
        if (tid == 32) {
            iscr[16] = 0; iscr[17] = -1; iscr[18] = 0; iscr[19] = 0; iscr[20] = 0; iscr[21] = 0; iscr[22] = 0;
            gap64[0] = ~0ull;
            sd[LS_GEN + 23] = nan("");                      // timestamp of the first valid point
        }
        __syncthreads();
        const FuseParams& prm = *reinterpret_cast<const FuseParams*>(sd + LS_PRM);
        const bool ekf_only = A.init_pos != nullptr;
        const double gap = prm.gap_threshold;

        // ------------------------------------------------------------------ phase A, sweep 1: flags, valid count, first gap
        // 256 poses per round (coalesced); per valid pose the previous valid pose comes from the warp ballot, the
        // warps before it in the round, or the carry of the earlier rounds.
        for (int lo = 0; lo < n; lo += LT) {
            const int i = lo + tid;
            bool v = false; double t = 0.0;
            if (i < n) {
                v = !row_has_nan(gz[3 * (size_t)i], gz[3 * (size_t)i + 1], gz[3 * (size_t)i + 2]);
                t = gts[i];
                gflag(i) = v ? FLAG_VALID : 0;
            }
            const unsigned bal = __ballot_sync(GSF_FULL_MASK, v);
            const int hi_l = bal ? 31 - __clz(bal) : 0, lo_l = bal ? __ffs(bal) - 1 : 0;
            const double w_last = __shfl_sync(GSF_FULL_MASK, t, hi_l), w_first = __shfl_sync(GSF_FULL_MASK, t, lo_l);
            const unsigned below = bal & ((1u << lane) - 1u);
            const int pl = below ? 31 - __clz(below) : 0;
            const double pt_w = __shfl_sync(GSF_FULL_MASK, t, pl);
            if (lane == 0) {
                iscr[warp] = __popc(bal); iscr[8 + warp] = lo + 32 * warp + hi_l;
                sd[LS_WARP + warp] = w_last; sd[LS_WARP + 8 + warp] = w_first;
            }
            __syncthreads();
            if (v) {
                double pT = 0.0; int pI = -1; bool have = false;
                if (below) { pT = pt_w; pI = lo + 32 * warp + pl; have = true; }
                else {
                    for (int w = warp - 1; w >= 0 && !have; --w)
                        if (iscr[w]) { pT = sd[LS_WARP + w]; pI = iscr[8 + w]; have = true; }
                    if (!have && iscr[16] > 0) { pT = sd[LS_GEN + 22]; pI = iscr[17]; have = true; }
                }
                // np.where(np.diff(ts[allv]) > gap)[0][0]: the first valid point that follows a gap, with its predecessor
                if (have && t - pT > gap) atomicMin(gap64, ((unsigned long long)(unsigned)i << 32) | (unsigned)pI);
            }
            __syncthreads();
            if (tid == 0) {
                int tot = 0;
                for (int w = 0; w < LNW; ++w) {
                    if (!iscr[w]) continue;
                    if (iscr[16] + tot == 0) sd[LS_GEN + 23] = sd[LS_WARP + 8 + w];
                    tot += iscr[w]; sd[LS_GEN + 22] = sd[LS_WARP + w]; iscr[17] = iscr[8 + w];
                }
                iscr[16] += tot;
            }
            __syncthreads();
        }
        const int nvalid = iscr[16];
        int st = ST_OK;

        // ------------------------------------------------------------------ recoveries (:879-894) and the sharp-turn gate
        for (int i = 1 + tid; i < n; i += LT) {
            if ((gflag(i) & FLAG_VALID) && !(gflag(i - 1) & FLAG_VALID)) {
                int s = i - 1;
                while (s > 0 && !(gflag(s - 1) & FLAG_VALID)) --s;
                int nf = FLAG_VALID | FLAG_RECOVERY;
                if (sharp_turn_ool(gts, A.quat + 4 * e0, s, i - 1, prm.yaw_rate_thresh)) nf |= FLAG_NO_RTS;
                gflag(i) = (unsigned char)nf;               // bit 0 is unchanged: concurrent walks over the flags stay valid
                iscr[19] = 1;
            }
        }

        // ------------------------------------------------------------------ phase A, sweeps 2-4: selection + centred sums
        if (!ekf_only) {
            // first = allv[:gaps[0]]: the slice ends at the gap index itself, i.e. BEFORE the last point of the first run
            const int i_cut = gap64[0] == ~0ull ? n : (int)(gap64[0] & 0xffffffffull);
            const double tlim = sd[LS_GEN + 23] + prm.max_duration;
            double c2[2] = {0.0, 0.0};
            for (int i = tid; i < i_cut; i += LT)
                if (gflag(i) & FLAG_VALID) { c2[0] += 1.0; if (gts[i] <= tlim) c2[1] += 1.0; }
            block_sum<2>(c2, sd + LS_SUMS);
            __syncthreads();
            const int first_cnt = (int)c2[0];
            const int mode = first_cnt < prm.min_samples ? 0 : ((int)c2[1] < prm.min_samples ? 1 : 2);
            double s7[7] = {0, 0, 0, 0, 0, 0, 0};
            for (int i = tid; i < n; i += LT) {
                const int f = gflag(i);
                if (!(f & FLAG_VALID)) continue;
                if (mode == 0 || (i < i_cut && (mode == 1 || gts[i] <= tlim))) {
                    gflag(i) = (unsigned char)(f | FLAG_SELECTED);
                    s7[0] += gpos[3 * (size_t)i]; s7[1] += gpos[3 * (size_t)i + 1]; s7[2] += gpos[3 * (size_t)i + 2];
                    s7[3] += gz[3 * (size_t)i]; s7[4] += gz[3 * (size_t)i + 1]; s7[5] += gz[3 * (size_t)i + 2];
                    s7[6] += 1.0;
                }
            }
            block_sum<7>(s7, sd + LS_SUMS);
            __syncthreads();
            const double inv_n = 1.0 / fmax(s7[6], 1.0);
            const double m0 = s7[0] * inv_n, m1 = s7[1] * inv_n, m2 = s7[2] * inv_n;
            const double d0 = s7[3] * inv_n, d1 = s7[4] * inv_n, d2 = s7[5] * inv_n;
            double h[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
            for (int i = tid; i < n; i += LT) {
                if (!(gflag(i) & FLAG_SELECTED)) continue;
                const double a0 = gpos[3 * (size_t)i] - m0, a1 = gpos[3 * (size_t)i + 1] - m1, a2 = gpos[3 * (size_t)i + 2] - m2;
                const double b0 = gz[3 * (size_t)i] - d0, b1 = gz[3 * (size_t)i + 1] - d1, b2 = gz[3 * (size_t)i + 2] - d2;
                h[0] += a0 * b0; h[1] += a0 * b1; h[2] += a0 * b2;
                h[3] += a1 * b0; h[4] += a1 * b1; h[5] += a1 * b2;
                h[6] += a2 * b0; h[7] += a2 * b1; h[8] += a2 * b2;
                h[9] += a0 * a0 + a1 * a1 + a2 * a2;
            }
            block_sum<10>(h, sd + LS_SUMS);
            __syncthreads();
            if (tid == 0) {
                double* g = sd + LS_GEN;
                g[0] = s7[6]; g[1] = m0; g[2] = m1; g[3] = m2; g[4] = d0; g[5] = d1; g[6] = d2;
#pragma unroll
                for (int k = 0; k < 10; ++k) g[7 + k] = h[k];
                iscr[21] = (int)s7[6]; iscr[22] = nvalid;
            }
        }
        __syncthreads();

        // ------------------------------------------------------------------ phase B: Umeyama finish, initial pose, C
        if (warp == 0) {
            int ust = 0;
            const Quat q0{A.quat[4 * e0], A.quat[4 * e0 + 1], A.quat[4 * e0 + 2], A.quat[4 * e0 + 3]};
            if (qnorm2(q0) == 0.0) ust |= ST_BAD_QUATERNION;
            if (!ekf_only) {
                const double* g = sd + LS_GEN;
                const int nsel = (int)g[0];
                double ms_[3], md_[3], hh[9], R[9], t[3], s = 1.0;
#pragma unroll
                for (int k = 0; k < 3; ++k) { ms_[k] = g[1 + k]; md_[k] = g[4 + k]; }
#pragma unroll
                for (int k = 0; k < 9; ++k) hh[k] = g[7 + k];
                if (nsel < 3 || nsel < prm.min_samples) ust |= ST_TOO_FEW_POINTS;
                else ust |= umeyama_finish_ool(nsel, ms_, md_, hh, g[16], R, t, &s);
                if (lane == 0 && !(ust & ST_TOO_FEW_POINTS)) {
                    const Quat qR = quat_from_matrix(R);
                    const Quat q0h = qunit(q0);
                    const Quat qs0 = qunit_or_identity(qmul(qR, q0h));
                    const Quat Cq = qmul(qs0, qconj(q0h));
                    double M[9]; qmat(Cq, M);
#pragma unroll
                    for (int k = 0; k < 9; ++k) { bc[k] = M[k]; bc[20 + k] = R[k]; }
                    bc[9] = Cq.x; bc[10] = Cq.y; bc[11] = Cq.z; bc[12] = Cq.w;
                    double rx, ry, rz;
                    mat_vec(R, gpos[0], gpos[1], gpos[2], rx, ry, rz);
                    bc[13] = s * rx + t[0]; bc[14] = s * ry + t[1]; bc[15] = s * rz + t[2];
                    bc[16] = t[0]; bc[17] = t[1]; bc[18] = t[2]; bc[19] = s;
                }
            } else if (lane == 0) {
                const Quat qi{A.init_quat[4 * b], A.init_quat[4 * b + 1], A.init_quat[4 * b + 2], A.init_quat[4 * b + 3]};
                const Quat qs0 = qunit_or_identity(qi);
                const Quat Cq = qmul(qs0, qconj(qunit(q0)));
                double M[9]; qmat(Cq, M);
#pragma unroll
                for (int k = 0; k < 9; ++k) bc[k] = M[k];
                bc[9] = Cq.x; bc[10] = Cq.y; bc[11] = Cq.z; bc[12] = Cq.w;
                bc[13] = A.init_pos[3 * b]; bc[14] = A.init_pos[3 * b + 1]; bc[15] = A.init_pos[3 * b + 2];
            }
            if (lane == 0) iscr[18] = ust;
        }
        __syncthreads();
        st |= iscr[18];
        if (st & (ST_TOO_FEW_POINTS | ST_BAD_QUATERNION)) {
            // The reference aborts the run here (ValueError / RuntimeError): outputs are NaN.
            for (long long i = tid; i < 3ll * n; i += LT) gout[i] = nan("");
            for (long long i = tid; i < 4ll * n; i += LT) gscr[i] = nan("");
            if (tid == 0) {
                A.status[b] = st;
                if (A.sim3_out) {
                    double* o = A.sim3_out + 16 * (size_t)b;
                    for (int k = 0; k < 13; ++k) o[k] = nan("");
                    o[13] = (double)iscr[21]; o[14] = (double)iscr[22]; o[15] = 0.0;
                }
            }
            continue;
        }

        // ------------------------------------------------------------------ phase C: EKF, tile by tile
        const bool xy_same = prm.p0[0] == prm.p0[1] && prm.q[0] == prm.q[1] && prm.r[0] == prm.r[1];
        const bool has_recovery = iscr[19] != 0;
        const double thr2 = (ekf_only || !(prm.residual_thresh > 0.0)) ? -1.0 : prm.residual_thresh * prm.residual_thresh;
        double w_sharp = 1.0;
        if (prm.sharp_turn_steps > 0) { const double wd = 1.0 / (double)prm.sharp_turn_steps; if (wd < 1.0) w_sharp = wd; }
        double* const tsS = ts_s + 1; double* const posS = pos_s + 3; double* const zS = z_s;
        unsigned char* const flgS = flg_s + 16;
        for (int i0 = 0; i0 < n; i0 += LTILE) {
            const int cnt = min(LTILE, n - i0);
            __syncthreads();                                // the previous tile is stored; carries are final
            for (int k = tid; k < cnt + 1; k += LT) {
                const int gi = i0 - 1 + k;
                ts_s[k] = gi >= 0 ? gts[gi] : 0.0;
                flg_s[15 + k] = gi >= 0 ? gflag(gi) : (unsigned char)0;
            }
            for (int k = tid; k < 3 * (cnt + 1); k += LT) { const long long g = 3ll * (i0 - 1) + k; pos_s[k] = g >= 0 ? gpos[g] : 0.0; }
            for (int k = tid; k < 3 * cnt; k += LT) z_s[k] = gz[3ll * i0 + k];
            __syncthreads();
            const int c0 = min(tid * LLCH, cnt), c1 = min(c0 + LLCH, cnt);
            const int s0 = i0 == 0 ? max(c0, 1) : c0;       // steps owned: local i in [s0, c1)
            double pprev0 = 0.0, pprev1 = 0.0, pprev2 = 0.0;
            if (s0 < c1) { pprev0 = posS[3 * (s0 - 1)]; pprev1 = posS[3 * (s0 - 1) + 1]; pprev2 = posS[3 * (s0 - 1) + 2]; }
            // covariance maps of the chunks, warp scan, warp totals
            const Moeb3 mex = xy_same ? moebius_chunk_scan<2>(tsS, flgS, prm, s0, c1, lane, sd + LS_MOEB + warp * 12)
                                      : moebius_chunk_scan<3>(tsS, flgS, prm, s0, c1, lane, sd + LS_MOEB + warp * 12);
            __syncthreads();
            Aff3 aex;
            double P[3];
            {
                Moeb3 pre = mex;
                if (warp > 0) {
                    Moeb3 acc;
#pragma unroll
                    for (int k = 0; k < 12; ++k) acc.m[k] = sd[LS_MOEB + k];
                    for (int w = 1; w < warp; ++w) {
                        Moeb3 nx;
#pragma unroll
                        for (int k = 0; k < 12; ++k) nx.m[k] = sd[LS_MOEB + w * 12 + k];
                        acc = moeb_compose<3>(acc, nx);
                    }
                    pre = moeb_compose<3>(acc, mex);
                }
#pragma unroll
                for (int a = 0; a < 3; ++a) {
                    const double* m = pre.m + 4 * a;
                    const double pin = i0 == 0 ? prm.p0[a] : sd[LS_CARRY + a];
                    P[a] = (m[0] * pin + m[1]) / (m[2] * pin + m[3]);
                }
                double RC[9];
#pragma unroll
                for (int k = 0; k < 9; ++k) RC[k] = bc[k];
                const double sc = bc[19], t0 = bc[16], t1 = bc[17], t2 = bc[18];
                int nviol = 0;
                Aff3 aff;
#pragma unroll
                for (int a = 0; a < 3; ++a) { aff.a[a] = 1.0; aff.b[a] = 0.0; }
                if (i0 == 0 && c0 == 0 && c1 > 0 && thr2 > 0.0 && (flgS[0] & FLAG_SELECTED)) {
                    double rx, ry, rz;
                    mat_vec(RC, posS[0], posS[1], posS[2], rx, ry, rz);
                    const double e0_ = sc * rx + t0 - zS[0], e1_ = sc * ry + t1 - zS[1], e2_ = sc * rz + t2 - zS[2];
                    if (!(e0_ * e0_ + e1_ * e1_ + e2_ * e2_ < thr2)) ++nviol;
                }
                double y0 = 0.0, y1 = 0.0, y2 = 0.0;        // Sim3 image of pose s0 - 1, advanced by s u each step
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
                    const int f = flgS[i];
                    const double qq[3] = {q0 * dt, q1 * dt, q2 * dt};
                    if (thr2 > 0.0) { y0 = fma(sc, u[0], y0); y1 = fma(sc, u[1], y1); y2 = fma(sc, u[2], y2); }
                    if (f & FLAG_VALID) {
                        const double zz[3] = {zS[3 * i], zS[3 * i + 1], zS[3 * i + 2]};
                        if (thr2 > 0.0 && (f & FLAG_SELECTED)) {
                            const double e0_ = y0 - zz[0], e1_ = y1 - zz[1], e2_ = y2 - zz[2];
                            if (!(e0_ * e0_ + e1_ * e1_ + e2_ * e2_ < thr2)) ++nviol;
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
                        if (f & FLAG_NO_RTS) {               // blended update at a sharp-turn recovery (:754-767)
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
#pragma unroll
                        for (int a = 0; a < 3; ++a) {
                            P[a] += qq[a];
                            posS[3 * i + a] = P[a];          // P_f[i] kept for the RTS patch
                            zS[3 * i + a] = u[a];
                            aff.b[a] += u[a];
                        }
                    }
                }
                if (i0 == 0 && c0 == 0 && c1 > 0) { posS[0] = prm.p0[0]; posS[1] = prm.p0[1]; posS[2] = prm.p0[2]; }   // P_f[0]
                if (thr2 > 0.0) {
                    nviol = warp_sum_i(nviol);
                    if (lane == 0 && nviol) atomicAdd(iscr + 20, nviol);
                }
                aff_warp_scan(aff, lane);
                if (lane == 31) {
#pragma unroll
                    for (int k = 0; k < 3; ++k) { sd[LS_AFF + warp * 6 + k] = aff.a[k]; sd[LS_AFF + warp * 6 + 3 + k] = aff.b[k]; }
                }
#pragma unroll
                for (int k = 0; k < 3; ++k) { aex.a[k] = __shfl_up_sync(GSF_FULL_MASK, aff.a[k], 1); aex.b[k] = __shfl_up_sync(GSF_FULL_MASK, aff.b[k], 1); }
                if (lane == 0) {
#pragma unroll
                    for (int k = 0; k < 3; ++k) { aex.a[k] = 1.0; aex.b[k] = 0.0; }
                }
            }
            __syncthreads();
            // state recursion from the carried state
            const bool last_owner = c0 < cnt && c1 == cnt;  // owns the last pose of the tile
            double x0, x1, x2;
            {
                Aff3 pre = aex;
                if (warp > 0) {
                    Aff3 acc;
#pragma unroll
                    for (int k = 0; k < 3; ++k) { acc.a[k] = sd[LS_AFF + k]; acc.b[k] = sd[LS_AFF + 3 + k]; }
                    for (int w = 1; w < warp; ++w) {
                        Aff3 nx;
#pragma unroll
                        for (int k = 0; k < 3; ++k) { nx.a[k] = sd[LS_AFF + w * 6 + k]; nx.b[k] = sd[LS_AFF + w * 6 + 3 + k]; }
                        acc = aff_compose(acc, nx);
                    }
                    pre = aff_compose(acc, aex);
                }
                const double xin0 = i0 == 0 ? bc[13] : sd[LS_CARRY + 3], xin1 = i0 == 0 ? bc[14] : sd[LS_CARRY + 4],
                             xin2 = i0 == 0 ? bc[15] : sd[LS_CARRY + 5];
                x0 = pre.a[0] * xin0 + pre.b[0]; x1 = pre.a[1] * xin1 + pre.b[1]; x2 = pre.a[2] * xin2 + pre.b[2];
                if (i0 == 0 && c0 == 0 && c1 > 0) { zS[0] = bc[13]; zS[1] = bc[14]; zS[2] = bc[15]; }
                for (int i = s0; i < c1; ++i) {
                    if (flgS[i] & FLAG_VALID) {
                        x0 = posS[3 * i] * x0 + zS[3 * i]; x1 = posS[3 * i + 1] * x1 + zS[3 * i + 1]; x2 = posS[3 * i + 2] * x2 + zS[3 * i + 2];
                    } else {
                        x0 += zS[3 * i]; x1 += zS[3 * i + 1]; x2 += zS[3 * i + 2];
                    }
                    zS[3 * i] = x0; zS[3 * i + 1] = x1; zS[3 * i + 2] = x2;
                }
            }
            __syncthreads();                                // every thread has read the carries of the previous tile
            if (last_owner) {
                sd[LS_CARRY] = P[0]; sd[LS_CARRY + 1] = P[1]; sd[LS_CARRY + 2] = P[2];
                sd[LS_CARRY + 3] = x0; sd[LS_CARRY + 4] = x1; sd[LS_CARRY + 5] = x2;
            }
            for (int k = tid; k < 3 * cnt; k += LT) gout[3ll * i0 + k] = zS[k];
            if (has_recovery) {                             // filtered covariances of the outage poses for the RTS patches
                for (int k = tid; k < 3 * cnt; k += LT) { const int j = k / 3; gscr[4ll * (i0 + j) + (k - 3 * j)] = posS[k]; }
            }
        }
        __syncthreads();

        // ------------------------------------------------------------------ phase D: closed-form RTS over recovered outages
        if (has_recovery) {
            for (int i = 1 + tid; i < n; i += LT) {
                const int f = gflag(i);
                if (!(f & FLAG_RECOVERY) || (f & FLAG_NO_RTS)) continue;
                int s = i - 1;
                while (s > 0 && !(gflag(s - 1) & FLAG_VALID)) --s;
                const double dt = fmax(1e-6, gts[i] - gts[i - 1]);
                const double* gp = gpos + 3 * (size_t)i;
                double u[3], ratio_den[3], delta[3];
                mat_vec(bc, gp[0] - gp[-3], gp[1] - gp[-2], gp[2] - gp[-1], u[0], u[1], u[2]);
#pragma unroll
                for (int a = 0; a < 3; ++a) {
                    ratio_den[a] = gscr[4ll * (i - 1) + a] + prm.q[a] * dt;                          // P_pred[i]
                    delta[a] = gout[3ll * i + a] - (gout[3ll * (i - 1) + a] + u[a]);                 // x_f[i] - x_pred[i]
                }
                for (int k = s; k < i; ++k) {
#pragma unroll
                    for (int a = 0; a < 3; ++a) gout[3ll * k + a] += (gscr[4ll * k + a] / ratio_den[a]) * delta[a];
                }
            }
        }
        __syncthreads();                                    // the scratch in out_quat is dead from here on

        // ------------------------------------------------------------------ phase E: quaternions, status
        if (iscr[20]) st |= ST_RANSAC_OUTLIERS;
        const Quat C{bc[9], bc[10], bc[11], bc[12]};
        const int bad = quat_rounds(A.quat + 4 * e0, A.out_quat + 4 * e0, C, tid, LT, n);
        if (tid == 0) {
            A.status[b] = st;
            if (A.sim3_out && !ekf_only) {
                double* o = A.sim3_out + 16 * (size_t)b;
#pragma unroll
                for (int k = 0; k < 9; ++k) o[k] = bc[20 + k];
                o[9] = bc[16]; o[10] = bc[17]; o[11] = bc[18]; o[12] = bc[19];
                o[13] = (double)iscr[21]; o[14] = (double)iscr[22]; o[15] = (double)iscr[20];
            }
        }
        __syncthreads();
        if (bad) atomicOr(A.status + b, ST_BAD_QUATERNION);
    }
}

cudaError_t launch_fuse_long(const FuseArgs& a, int num_sms, cudaStream_t stream) {
    const size_t smem = long_smem_bytes();
    cudaError_t e = cudaFuncSetAttribute(fuse_long_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    long long grid = num_sms;
    if (grid > a.B) grid = a.B;
    fuse_long_kernel<<<(unsigned)grid, LT, smem, stream>>>(a);
    return cudaGetLastError();
}

}  // namespace gsf
