// K1b for ONE long trajectory: dynamic_time_alignment (/root/reference/EKFGPSSLAM.py:325-387) when the GNSS track has
// millions of samples (BASELINE config 4: 1e8 poses).  The per-trajectory kernel of gsf_kernels.cu solves the not-a-knot
// spline of a segment with one serial Thomas sweep per axis -- fine for 271 knots, hopeless for 1e8.  Here the solve is
// LOCAL: the spline's tridiagonal system is diagonally dominant (|off-diagonal| <= diagonal / 2 for any knot spacing), so the
// influence of a boundary value on the moment k knots away decays at least like (2 - sqrt 3)^k = 0.268^k.  Every thread
// owns AL_CH = 13 consecutive knots and solves the system on its chunk widened by a halo of AL_H = 20 knots on both sides,
// with a natural end (m = 0) where the halo cuts the segment and the true not-a-knot rows where the segment really ends
// inside the halo; only the chunk's own moments are kept.  The cut moment is wrong by its own size, and 0.268^20 = 3.7e-12
// of that reaches the chunk: for GNSS noise of 0.3 m at 10 Hz (second differences of 30 m/s^2, a curvature term of
// m h^2 / 8 = 4 cm in the interpolant) that is 1.5e-13 m, far below one ulp of a UTM coordinate (9.3e-10 m) -- the values
// equal the global solve's (the test compares with scipy and with the serial kernel).
//
// Layout of the moments kernel: a block of 128 threads stages its window of 1664 + 2 x 22 knots (t and xyz, exactly as they
// lie in global memory: 13- and 39-double strides between threads are conflict-free) with two TMA bulk copies, so every knot
// comes from DRAM once (the halo twice).  A thread eliminates forward from the left end of its window to the last knot of its
// chunk, keeping the eliminated rows of its own 13 knots in REGISTERS (fully unrolled, predicated on the segment limits),
// eliminates from the right end of the window down to the knot after its chunk keeping nothing, joins the two at the chunk's
// last row (a 2 x 2 system) and back-substitutes through its registers.  (First version: 64-knot chunks with the whole
// widened system in local memory -- 9.8 ms for 1e8 knots, 4.2x the algorithmic DRAM traffic from local-memory spills; this one:
// 1.7 ms.)  Only the INTERIOR moments of a segment are stored; the two end moments follow from the not-a-knot rows
// (m_a = ((h0 + h1) m_{a+1} - h0 m_{a+2}) / h1) and are formed by the evaluation when a stamp falls into a segment's first or
// last interval.  Segments split at gaps > max_gps_gap_threshold (:351-354); 2-3 knot segments are linear (:362), single
// knots give nothing (:361).  A block whose window holds no gap and no end of the track (the usual case) skips the walk to
// the segment ends.
//
// Evaluation: a tile of 2048 stamps per block.  assoc_long_bracket_kernel (one warp per tile) brackets the tile's smallest and
// largest stamp in the knot array with a 32-way search (6 dependent reads for 1e8 knots); the evaluation kernel stages the
// bracket's knot times in shared memory, finds each stamp's interval from the position equally spaced knots would give
// (verified, else bisection of the bracket), decides segment membership from the gaps around it, and evaluates the cubic in
// moment form / the line / NaN (scipy: NaN outside [seg start, seg end], :377-379).
#include "gsf_common.cuh"
#include "gsf_fuse_shared.cuh"
#include "gsf_internal.cuh"

namespace gsf {

#ifndef GSF_AL_CH
#define GSF_AL_CH 13                  // 13: 1.71 ms per 1e8 knots; 15: 1.87 (register spills at 168); 17 with two blocks per SM: 1.71
#endif
#ifndef GSF_AL_MINB
#define GSF_AL_MINB 3
#endif
#ifndef GSF_EV_PER
#define GSF_EV_PER 8
#endif
#ifndef GSF_EV_MINB
#define GSF_EV_MINB 4
#endif
constexpr int AL_CH = GSF_AL_CH;                  // knots per thread (odd: conflict-free shared-memory strides of AL_CH and 3 AL_CH doubles)
constexpr int AL_H = 20;                          // halo
constexpr int AL_NT = 128;
constexpr int AL_TILE = AL_CH * AL_NT;            // knots per block
constexpr int AL_PADW = AL_H + 2;                 // window margin (even, so that the window starts on a 16-byte boundary)
constexpr int AL_W = AL_TILE + 2 * AL_PADW;
constexpr size_t AL_SMEM = (size_t)AL_W * 32 + 16;

__global__ void __launch_bounds__(AL_NT, GSF_AL_MINB) assoc_long_moments_kernel(const double* __restrict__ gt, const double* __restrict__ gy, long long M, double gap,
                                                                      double* __restrict__ mom, int* __restrict__ bad_steps, int use_tma) {
    extern __shared__ __align__(16) double al_sm[];
    double* Yv = al_sm;                           // [AL_W][3]
    double* T = al_sm + 3 * AL_W;                 // [AL_W]
    uint64_t* mbar = reinterpret_cast<uint64_t*>(T + AL_W);
    const int tid = threadIdx.x;
    const long long K0 = (long long)blockIdx.x * AL_TILE, base = K0 - AL_PADW;
    // ---- the block's window of knots -> shared memory (every knot is read from DRAM once; the halo twice)
    if (use_tma && base >= 0 && base + AL_W <= M) {
        if (tid == 0) { mbar_init(mbar, 1); fence_mbar_init(); }
        __syncthreads();
        if (tid == 0) {
            mbar_expect_tx(mbar, (uint32_t)AL_W * 32u);
            bulk_g2s(T, gt + base, (uint32_t)AL_W * 8u, mbar);
            bulk_g2s(Yv, gy + 3 * base, (uint32_t)AL_W * 24u, mbar);
        }
        mbar_wait(mbar, 0);
    } else {
        const long long jlo = max(base, 0ll), jhi = min(base + AL_W, M);
        for (long long j = jlo + tid; j < jhi; j += AL_NT) T[j - base] = gt[j];
        for (long long f = 3 * jlo + tid; f < 3 * jhi; f += AL_NT) Yv[f - 3 * base] = gy[f];
        __syncthreads();
    }
    const int kfirst = base <= 0 ? (int)(-base) : -1;                          // local index of knot 0 / knot M-1 when inside the window
    const int klast = (M - 1 - base < AL_W) ? (int)(M - 1 - base) : (1 << 30);
    const int kc0 = AL_PADW + AL_CH * tid;                                    // the thread's chunk, local indices
    const int kj1 = min(kc0 + AL_CH - 1, klast);
    double cpR[AL_CH], dR[3][AL_CH];
#pragma unroll
    for (int L = 0; L < AL_CH; ++L) { cpR[L] = 0.0; dR[0][L] = 0.0; dR[1][L] = 0.0; dR[2][L] = 0.0; }
    // stamps that do not increase by more than 1e-9 inside a segment: the reference drops the segment (:356-359); and one
    // block-wide test for the usual case -- no gap and no end of the track anywhere in the window -- which spares every
    // thread the walk to its segment's ends
    int gap_seen = (kfirst >= 0) || (klast < AL_W);
    for (int k = kc0; k <= kj1 && k < klast; ++k) {
        const double st = T[k + 1] - T[k];
        if (st > gap) gap_seen = 1; else if (!(st > 1e-9)) *bad_steps = 1;
    }
    if (!gap_seen && tid < 2 * AL_PADW - 1) {                 // the margins of the window
        const int k = tid < AL_PADW ? tid : AL_TILE + tid;
        if (T[k + 1] - T[k] > gap) gap_seen = 1;
    }
    const bool plain = !__syncthreads_or(gap_seen);
    int cur = kc0;
    while (cur <= kj1) {
        // segment extent around `cur`, limited to the halo
        int a = cur, b = cur; bool lt = false, rt = false;
        if (plain) { a = kc0 - AL_H; b = kj1 + AL_H; }
        else {
            for (;;) {
                if (a == kfirst) { lt = true; break; }
                if (T[a] - T[a - 1] > gap) { lt = true; break; }
                if (a <= kc0 - AL_H) break;
                --a;
            }
            for (;;) {
                if (b == klast) { rt = true; break; }
                if (T[b + 1] - T[b] > gap) { rt = true; break; }
                if (b - kj1 >= AL_H) break;
                ++b;
            }
        }
        const int w1 = min(b, kj1);
        // interior unknowns U0 .. U1 (the end moments follow from the not-a-knot rows at evaluation time); this chunk's: R0 .. R1
        const int U0 = a + 1, U1 = b - 1;
        const int R0 = max(max(a, kc0), U0), R1 = min(w1, U1);
        if (b - a + 1 >= 4 && R0 <= R1) {          // fewer than 4 knots (both ends true then): linear, no moments (:362)
            const double e_h0 = T[a + 1] - T[a], e_h1 = T[a + 2] - T[a + 1], e_ha = T[b - 1] - T[b - 2], e_hb = T[b] - T[b - 1];
            const double l_di = lt ? e_h0 * (e_h0 + e_h1) / e_h1 : 0.0, l_up = lt ? e_h0 * e_h0 / e_h1 : 0.0;     // not-a-knot rows at true ends
            const double r_di = rt ? e_hb * (e_ha + e_hb) / e_ha : 0.0, r_lo = rt ? e_hb * e_hb / e_ha : 0.0;     // (natural where the halo cuts)
            // ---- forward elimination U0 .. R1: knot times, spacings and slopes carried from row to row
            double pcp = 0.0, pd[3] = {0.0, 0.0, 0.0};
            double tj = T[a + 1], hp = e_h0;
            double yj[3], sp[3];
            {
                const double rh = fast_rcp(hp);
#pragma unroll
                for (int ax = 0; ax < 3; ++ax) { yj[ax] = Yv[3 * (a + 1) + ax]; sp[ax] = (yj[ax] - Yv[3 * a + ax]) * rh; }
            }
            auto fstep = [&](int j, double& c_out, double* d_out) {
                const double tn = T[j + 1], hc = tn - tj;
                double lo = hp, di = 2.0 * (hp + hc), up = hc;
                if (j == U0) { di += l_di; up -= l_up; lo = 0.0; }
                if (j == U1) { di += r_di; lo -= r_lo; up = 0.0; }
                const double r = fast_rcp(di - lo * pcp), rh = fast_rcp(hc);
                c_out = up * r;
#pragma unroll
                for (int ax = 0; ax < 3; ++ax) {
                    const double yn = Yv[3 * (j + 1) + ax];
                    const double sn = (yn - yj[ax]) * rh;
                    const double d = (6.0 * (sn - sp[ax]) - lo * pd[ax]) * r;
                    d_out[ax] = d; pd[ax] = d; yj[ax] = yn; sp[ax] = sn;
                }
                pcp = c_out; tj = tn; hp = hc;
            };
#pragma unroll 4
            for (int j = U0; j < kc0; ++j) { double c, d[3]; fstep(j, c, d); }
#pragma unroll
            for (int L = 0; L < AL_CH; ++L) {
                const int j = kc0 + L;
                if (j >= R0 && j <= R1) { double c, d[3]; fstep(j, c, d); cpR[L] = c; dR[0][L] = d[0]; dR[1][L] = d[1]; dR[2][L] = d[2]; }
            }
            // ---- elimination from the right end U1 .. R1 + 1 (nothing stored), then the 2 x 2 system that joins the two at row R1
            double pbp = 0.0, pe[3] = {0.0, 0.0, 0.0};
            if (U1 > R1) {
                double tb = T[b - 1], hn = e_hb;
                double yb[3], sn[3];
                {
                    const double rh = fast_rcp(hn);
#pragma unroll
                    for (int ax = 0; ax < 3; ++ax) { yb[ax] = Yv[3 * (b - 1) + ax]; sn[ax] = (Yv[3 * b + ax] - yb[ax]) * rh; }
                }
#pragma unroll 4
                for (int j = U1; j > R1; --j) {
                    const double tp = T[j - 1], hq = tb - tp;
                    double lo = hq, di = 2.0 * (hq + hn), up = hn;
                    if (j == U1) { di += r_di; lo -= r_lo; up = 0.0; }
                    const double r = fast_rcp(di - up * pbp), rh = fast_rcp(hq);
#pragma unroll
                    for (int ax = 0; ax < 3; ++ax) {
                        const double yp = Yv[3 * (j - 1) + ax];
                        const double sq = (yb[ax] - yp) * rh;
                        const double e = (6.0 * (sn[ax] - sq) - up * pe[ax]) * r;
                        pe[ax] = e; yb[ax] = yp; sn[ax] = sq;
                    }
                    pbp = lo * r; tb = tp; hn = hq;
                }
            }
            double mn[3] = {0.0, 0.0, 0.0};
#pragma unroll
            for (int L = AL_CH - 1; L >= 0; --L) {
                const int j = kc0 + L;
                if (j >= R0 && j <= R1) {
                    if (j == R1) {
                        const double inv = U1 > R1 ? fast_rcp(1.0 - cpR[L] * pbp) : 1.0;
#pragma unroll
                        for (int ax = 0; ax < 3; ++ax) mn[ax] = (dR[ax][L] - cpR[L] * pe[ax]) * inv;
                    } else {
#pragma unroll
                        for (int ax = 0; ax < 3; ++ax) mn[ax] = dR[ax][L] - cpR[L] * mn[ax];
                    }
                    dR[0][L] = mn[0]; dR[1][L] = mn[1]; dR[2][L] = mn[2];
                }
            }
        }
        cur = w1 + 1;
    }
    // ---- moments -> shared memory (over the y window, which nobody reads any more) -> coalesced store
    __syncthreads();
#pragma unroll
    for (int L = 0; L < AL_CH; ++L) { Yv[3 * (kc0 + L)] = dR[0][L]; Yv[3 * (kc0 + L) + 1] = dR[1][L]; Yv[3 * (kc0 + L) + 2] = dR[2][L]; }
    __syncthreads();
    const long long cnt = 3 * min((long long)AL_TILE, M - K0);
    for (long long f = tid; f < cnt; f += AL_NT) mom[3 * K0 + f] = Yv[3 * AL_PADW + f];
}

// largest j with gt[j] <= q (0 if none) and the smallest bracket end above it, found by one warp probing 32 knots per round
// (6 dependent rounds for 1e8 knots instead of the 27 of a bisection)
__device__ __forceinline__ void warp_bracket(const double* __restrict__ gt, long long M, double q, int lane, long long& l_out, long long& r_out) {
    long long l = 0, r = M - 1;
    while (r - l > 1) {
        const long long pos = l + ((r - l) * (lane + 1)) / 33;
        const bool le = pos == l || gt[pos] <= q;
        const int n = __popc(__ballot_sync(GSF_FULL_MASK, le));             // the knots are sorted: the lanes with le come first
        const long long nl = __shfl_sync(GSF_FULL_MASK, pos, max(n - 1, 0)), nr = __shfl_sync(GSF_FULL_MASK, pos, min(n, 31));
        if (n < 32) r = nr;
        if (n > 0) l = nl;
    }
    l_out = l; r_out = r;
}

constexpr int EV_NT = 256;
constexpr int EV_PER = GSF_EV_PER;                        // stamps per thread: one bracket search serves EV_NT * EV_PER stamps
constexpr int EV_TILE = EV_NT * EV_PER;
constexpr int EV_CAP = EV_TILE + EV_TILE / 2;                     // knot times of the block's bracket staged in shared memory
constexpr int EV_MARGIN = 3;                     // knots around the bracket that the segment tests read

// Interpolant on the interval [j, j+1] (knot times tj, tn) of a segment that has nl / nr (0..2) more knots on the left / right;
// hl, hr: the neighbouring spacings when those knots exist.
__device__ __forceinline__ void eval_interval(double t, double tj, double tn, double hl, double hr, int nl, int nr, long long j,
                                              const double* __restrict__ gy, const double* __restrict__ mom, double& v0, double& v1, double& v2) {
    const double hh = tn - tj;
    const double* __restrict__ y = gy + 3 * j;
    double v[3];
    if (2 + nl + nr >= 4) {
        const double* __restrict__ m = mom + 3 * j;
        const double wb = (t - tj) / hh, wa = 1.0 - wb;
        const double h26 = hh * hh * (1.0 / 6.0);
        const double ca = (wa * wa * wa - wa) * h26, cb = (wb * wb * wb - wb) * h26;
#pragma unroll
        for (int ax = 0; ax < 3; ++ax) {
            // the moments kernel leaves the interior moments; the two end moments of a segment come from its not-a-knot rows
            const double m0 = nl ? m[ax] : ((hh + hr) * m[3 + ax] - hh * m[6 + ax]) / hr;
            const double m1 = nr ? m[3 + ax] : ((hl + hh) * m[ax] - hh * m[ax - 3]) / hl;
            v[ax] = fma(wb, y[3 + ax] - y[ax], y[ax]) + (ca * m0 + cb * m1);
        }
    } else {
#pragma unroll
        for (int ax = 0; ax < 3; ++ax) v[ax] = (y[3 + ax] - y[ax]) / hh * (t - tj) + y[ax];          // 2-3 knots: linear (:362)
    }
    v0 = v[0]; v1 = v[1]; v2 = v[2];
}

// One stamp against the whole knot array (brackets too wide to stage).
__device__ __forceinline__ void eval_stamp_global(double t, long long bl, long long br, const double* __restrict__ gt, const double* __restrict__ gy,
                                                  const double* __restrict__ mom, long long M, double gap, double t_first, double t_last,
                                                  double& v0, double& v1, double& v2) {
    v0 = v1 = v2 = nan("");
    if (!(t >= t_first && t <= t_last)) return;
    long long l = bl, r = br;                                 // largest j with gt[j] <= t: gt[b_l] <= t (or b_l = 0), gt[b_r] >= t (or b_r = M - 1)
    while (r - l > 1) { const long long mid = (l + r) >> 1; if (gt[mid] <= t) l = mid; else r = mid; }
    long long j = (gt[r] <= t) ? r : l;
    // interval [j, j+1] unless t sits exactly on the last knot of a segment: then [j-1, j]
    double tj = gt[j];
    if (j == M - 1 || gt[j + 1] - tj > gap) {
        if (t == tj && j > 0 && !(tj - gt[j - 1] > gap)) { --j; tj = gt[j]; } else return;        // inside a gap (or an isolated knot): no segment
    }
    const double tn = gt[j + 1];
    double hl = 0.0, hr = 0.0;
    int nl = 0, nr = 0;
    if (j >= 1) { const double tp = gt[j - 1]; hl = tj - tp; if (!(hl > gap)) { nl = 1; if (j >= 2 && !(tp - gt[j - 2] > gap)) nl = 2; } }
    if (j + 2 <= M - 1) { const double tq = gt[j + 2]; hr = tq - tn; if (!(hr > gap)) { nr = 1; if (j + 3 <= M - 1 && !(gt[j + 3] - tq > gap)) nr = 2; } }
    eval_interval(t, tj, tn, hl, hr, nl, nr, j, gy, mom, v0, v1, v2);
}

// One stamp against the staged bracket: S[k] = gt[s0 + k]; lo / hi: the bracket ends; zl: 0 if s0 == 0 (knot 0 is S[0]) else a
// negative number (every S[jl - 2] exists); ml: index of knot M - 1 if it is staged, else huge.  The search starts from the
// position that equally spaced knots would give and falls back to the bisection of the bracket when that misses.
__device__ __forceinline__ void eval_stamp_staged(double t, const double* __restrict__ S, int lo, int hi, double inv_span, long long s0, int zl, int ml,
                                                  const double* __restrict__ gy, const double* __restrict__ mom, double gap, double t_first,
                                                  double t_last, double& v0, double& v1, double& v2) {
    v0 = v1 = v2 = nan("");
    if (!(t >= t_first && t <= t_last)) return;
    int g = lo + (int)((t - S[lo]) * inv_span);
    g = max(lo, min(g, hi));
    const int wl = max(g - 3, lo), wr = min(g + 4, hi);
    int li = lo, ri = hi;
    if ((wl == lo || S[wl] <= t) && (wr == hi || S[wr] > t)) { li = wl; ri = wr; }
    while (ri - li > 1) { const int mid = (li + ri) >> 1; if (S[mid] <= t) li = mid; else ri = mid; }
    int jl = (S[ri] <= t) ? ri : li;
    double tj = S[jl];
    if (jl == ml || S[jl + 1] - tj > gap) {
        if (t == tj && jl >= 1 + zl && !(tj - S[jl - 1] > gap)) { --jl; tj = S[jl]; } else return;
    }
    const double tn = S[jl + 1];
    double hl = 0.0, hr = 0.0;
    int nl = 0, nr = 0;
    if (jl >= 1 + zl) { const double tp = S[jl - 1]; hl = tj - tp; if (!(hl > gap)) { nl = 1; if (jl >= 2 + zl && !(tp - S[jl - 2] > gap)) nl = 2; } }
    if (jl + 2 <= ml) { const double tq = S[jl + 2]; hr = tq - tn; if (!(hr > gap)) { nr = 1; if (jl + 3 <= ml && !(S[jl + 3] - tq > gap)) nr = 2; } }
    eval_interval(t, tj, tn, hl, hr, nl, nr, s0 + jl, gy, mom, v0, v1, v2);
}

// One warp per tile of EV_TILE stamps: the knot interval that brackets the tile's smallest and largest stamp.  A chain of
// dependent DRAM reads per tile -- in its own kernel, where ten thousand warps wait side by side, not in front of the evaluation.
__global__ void __launch_bounds__(256) assoc_long_bracket_kernel(const double* __restrict__ gt, long long M, const double* __restrict__ st, long long N,
                                                                 long long* __restrict__ brackets, long long ntiles) {
    const long long w = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (w >= ntiles) return;
    double mn = INFINITY, mx = -INFINITY;
    const long long i0 = w * EV_TILE, i1 = min(i0 + EV_TILE, N);
#pragma unroll 8
    for (long long i = i0 + lane; i < i1; i += 32) { const double t = st[i]; if (t == t) { mn = fmin(mn, t); mx = fmax(mx, t); } }
    for (int o = 16; o > 0; o >>= 1) { mn = fmin(mn, __shfl_xor_sync(GSF_FULL_MASK, mn, o)); mx = fmax(mx, __shfl_xor_sync(GSF_FULL_MASK, mx, o)); }
    long long l, r, l2, r2;
    warp_bracket(gt, M, mn, lane, l, r);
    warp_bracket(gt, M, mx, lane, l2, r2);
    if (lane == 0) { brackets[2 * w] = l; brackets[2 * w + 1] = r2; }
}

__global__ void __launch_bounds__(EV_NT, GSF_EV_MINB) assoc_long_eval_kernel(const double* __restrict__ gt, const double* __restrict__ gy, const double* __restrict__ mom,
                                                                long long M, const double* __restrict__ st, long long N, double gap,
                                                                const long long* __restrict__ brackets, double* __restrict__ out,
                                                                unsigned char* __restrict__ val) {
    // The stamps of a block are usually close together (SLAM stamps are sorted): assoc_long_bracket_kernel has bracketed the
    // block's smallest and largest stamp in the whole knot array; the knot times of the bracket go to shared memory (when it is
    // small enough) and every thread searches there.
    __shared__ double s_t[EV_CAP];
    const long long i0 = (long long)blockIdx.x * EV_TILE + threadIdx.x;
    if (M < 2) {
#pragma unroll
        for (int u = 0; u < EV_PER; ++u) {
            const long long i = i0 + (long long)u * EV_NT;
            if (i < N) { out[3 * i] = nan(""); out[3 * i + 1] = nan(""); out[3 * i + 2] = nan(""); val[i] = 0; }
        }
        return;
    }
    const long long bl = brackets[2 * blockIdx.x], br = brackets[2 * blockIdx.x + 1];
    const long long s0 = max(bl - EV_MARGIN, 0ll), s1 = min(br + EV_MARGIN, M - 1);
    const bool staged = br >= bl && s1 - s0 < EV_CAP;
    if (staged) {
        for (long long k = s0 + threadIdx.x; k <= s1; k += EV_NT) s_t[k - s0] = gt[k];
        __syncthreads();
    }
    const double t_first = gt[0], t_last = gt[M - 1];
    const int lo = (int)(bl - s0), hi = (int)(br - s0);
    const int zl = s0 == 0 ? 0 : -8, ml = s1 == M - 1 ? (int)(s1 - s0) : (1 << 30);
    double inv_span = 0.0;
    if (staged) { const double span = s_t[hi] - s_t[lo]; inv_span = span > 0.0 ? (double)(hi - lo) / span : 0.0; }
#pragma unroll 2
    for (int u = 0; u < EV_PER; ++u) {
        const long long i = i0 + (long long)u * EV_NT;
        if (i >= N) break;
        const double t = st[i];
        double v0, v1, v2;
        if (staged) eval_stamp_staged(t, s_t, lo, hi, inv_span, s0, zl, ml, gy, mom, gap, t_first, t_last, v0, v1, v2);
        else eval_stamp_global(t, bl, br, gt, gy, mom, M, gap, t_first, t_last, v0, v1, v2);
        out[3 * i] = v0; out[3 * i + 1] = v1; out[3 * i + 2] = v2;
        val[i] = !row_has_nan(v0, v1, v2);
    }
}

// work: 3 M doubles (moments), 1 int (+ padding), 2 int64 per tile of EV_TILE stamps (brackets)
long long associate_long_work_doubles(long long M, long long N) { return 3 * M + 2 + 2 * ((N + EV_TILE - 1) / EV_TILE); }

cudaError_t launch_associate_long(const double* gps_t, const double* gps_xyz, long long M, const double* slam_t, long long N, double gap,
                                  double* work, double* aligned, unsigned char* valid, int* status, cudaStream_t stream) {
    int* bad = reinterpret_cast<int*>(work + 3 * M);
    long long* brackets = reinterpret_cast<long long*>(work + 3 * M + 2);
    cudaError_t e = cudaMemsetAsync(bad, 0, sizeof(int), stream);
    if (e != cudaSuccess) return e;
    if (M >= 2) {
        static bool attr_done = false;
        if (!attr_done) {
            e = cudaFuncSetAttribute(assoc_long_moments_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)AL_SMEM);
            if (e != cudaSuccess) return e;
            attr_done = true;
        }
        const int use_tma = ((reinterpret_cast<uintptr_t>(gps_t) | reinterpret_cast<uintptr_t>(gps_xyz)) & 15) == 0;
        assoc_long_moments_kernel<<<(unsigned)((M + AL_TILE - 1) / AL_TILE), AL_NT, AL_SMEM, stream>>>(gps_t, gps_xyz, M, gap, work, bad, use_tma);
    }
    if (N > 0) {
        const long long ntiles = (N + EV_TILE - 1) / EV_TILE;
        if (M >= 2) assoc_long_bracket_kernel<<<(unsigned)((ntiles + 7) / 8), 256, 0, stream>>>(gps_t, M, slam_t, N, brackets, ntiles);
        assoc_long_eval_kernel<<<(unsigned)ntiles, EV_NT, 0, stream>>>(gps_t, gps_xyz, work, M, slam_t, N, gap, brackets, aligned, valid);
    }
    e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    if (status) e = cudaMemcpyAsync(status, bad, sizeof(int), cudaMemcpyDeviceToDevice, stream);
    return e;
}

}  // namespace gsf
