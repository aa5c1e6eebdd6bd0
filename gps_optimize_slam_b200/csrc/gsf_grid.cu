// gsf_ekf_hypothesis_grid: ONE trajectory, H hypotheses of the EKF noise parameters (P0, Q, R diagonals);
// per hypothesis the EKF of apply_ekf_correction (EKFGPSSLAM.py:831-935, ExtendedKalmanFilter :679-772) is
// run from the Sim3-aligned pose 0 and scored with the nearest-neighbour error statistics of :1021-1033
// (mean / median / RMSE).  BASELINE.json config 5 (KITTI-00 length x 262 144 noise-grid hypotheses).
//
// Everything that does not depend on the hypothesis is computed once (prep kernels): Sim3 selection
// (:972-998), Umeyama (:428-459, tiled kernels of gsf_kernels.cu), C = q_state0 (x) conj(q_hat0), the
// telescoped odometry u_i = M(C)(p_i - p_{i-1}) (see gsf_fused.cu), the step records {dt, u, z} and the
// candidate set sorted by x with every evaluation pose's own rank in it.
// The grid kernel maps one THREAD to one hypothesis: state (x, P per axis) and the nine noise values stay
// in registers, the shared step records are staged through shared memory with TMA bulk copies
// (double-buffered tiles, mbarrier), the candidate set is resident in shared memory.  The nearest
// neighbour is found exactly by scanning outwards in the x-sorted candidate order from the pose's own
// measurement until the 1-D gap alone exceeds the best distance (the pose's own measurement is always a
// candidate, so the first bound is its innovation).  Errors go to a [poses, hypotheses] table (coalesced);
// a second kernel radix-selects the median of each hypothesis' column.  FP64-pipe bound; no tensor cores.
// Restriction (status GSF_ST_GRID_NEEDS_ALL_VALID otherwise): every pose has a GNSS measurement, i.e. no
// outage / RTS segments (those need per-hypothesis history; use gsf_fuse_batched_dev with per-trajectory
// parameters for such data).
#include <cstdio>
#include <cstdlib>
#include "gsf_common.cuh"
#include "gsf_ptx.cuh"
#include "gsf_internal.cuh"
#include "gsf_select.cuh"

namespace gsf {

constexpr int GRID_REC = 8;          // doubles per step record: dt, u(3), z(3), own rank (or -1: not evaluated)
constexpr int GRID_TILE = 128;       // step records per shared-memory tile (8 KB)
// Hypotheses per block are a template parameter: the candidate set takes ~110 KB of shared memory, so one block
// is resident per SM and its thread count IS the occupancy (256 -> 83 ms, 512 -> 66 ms, 1024 -> 62 ms on config 5);
// small grids use smaller blocks so that every SM gets one.
constexpr int GRID_FATAL = ST_TOO_FEW_POINTS | ST_BAD_QUATERNION | ST_GRID_NEEDS_ALL_VALID;

// ----------------------------------------------------------------------------- prep 1: validity, Sim3 selection mask
__global__ void __launch_bounds__(1024) grid_prep_select_kernel(const double* __restrict__ ts, const double* __restrict__ z, int n,
                                                                const FuseParams* __restrict__ prm, unsigned char* __restrict__ mask,
                                                                long long* __restrict__ offsets2, int* __restrict__ status) {
    __shared__ int s_invalid, s_gap0, s_timed;
    if (threadIdx.x == 0) { s_invalid = 0; s_gap0 = 0x7fffffff; s_timed = 0; }
    __syncthreads();
    const double gap = prm->gap_threshold;
    for (int i = threadIdx.x; i < n; i += 1024) {
        if (row_has_nan(z[3 * i], z[3 * i + 1], z[3 * i + 2])) s_invalid = 1;
        if (i + 1 < n && ts[i + 1] - ts[i] > gap) atomicMin(&s_gap0, i);      // np.where(np.diff(ts[allv]) > gap)[0][0]
    }
    __syncthreads();
    // :977-998 with every pose valid: first = allv[:gaps[0]] (the reference's slice end is the gap index itself)
    const int first_cnt = s_gap0 == 0x7fffffff ? n : s_gap0;
    const double tlim = ts[0] + prm->max_duration;
    int mode = 0;                                           // 0: all, 1: first run, 2: first run within max_duration
    if (first_cnt >= prm->min_samples) {
        int c = 0;
        for (int i = threadIdx.x; i < first_cnt; i += 1024) c += ts[i] <= tlim ? 1 : 0;
        atomicAdd(&s_timed, c);
        __syncthreads();
        mode = s_timed < prm->min_samples ? 1 : 2;
    }
    for (int i = threadIdx.x; i < n; i += 1024)
        mask[i] = (mode == 0 || (i < first_cnt && (mode == 1 || ts[i] <= tlim))) ? 1 : 0;
    if (threadIdx.x == 0) {
        const int nsel = mode == 0 ? n : (mode == 1 ? first_cnt : s_timed);
        int st = s_invalid ? ST_GRID_NEEDS_ALL_VALID : 0;
        if (nsel < 3 || nsel < prm->min_samples) st |= ST_TOO_FEW_POINTS;       // ValueError :975 / :997
        offsets2[0] = 0; offsets2[1] = n; status[0] = st;
    }
}

// ----------------------------------------------------------------------------- prep 2: step records + x-sorted candidates
// One block.  rec[i] = {dt_i, u_i, z_i, rank_i}; cand[k] = k-th evaluation measurement in x order;
// hdr = {x0(3), n_eval}.
__global__ void __launch_bounds__(1024) grid_prep_records_kernel(const double* __restrict__ ts, const double* __restrict__ pos,
                                                                 const double* __restrict__ quat, const double* __restrict__ z, int n,
                                                                 const FuseParams* __restrict__ prm, const double* __restrict__ R,
                                                                 const double* __restrict__ t, const double* __restrict__ s,
                                                                 double* __restrict__ rec, double* __restrict__ cand, double* __restrict__ hdr,
                                                                 int* __restrict__ status, int cap_pow2) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double* key = reinterpret_cast<double*>(smem_raw);      // [cap_pow2] x of the evaluation measurements
    int* idx = reinterpret_cast<int*>(key + cap_pow2);      // [cap_pow2] pose index
    __shared__ double MC[9];
    __shared__ int s_count;
    if (threadIdx.x == 0) {
        s_count = 0;
        const Quat q0{quat[0], quat[1], quat[2], quat[3]};
        int st = status[0] | status[1];                     // selection status | Umeyama status
        if (qnorm2(q0) == 0.0) st |= ST_BAD_QUATERNION;
        const Quat qR = quat_from_matrix(R);
        const Quat q0h = qunit(q0);
        const Quat qs0 = qunit_or_identity(qmul(qR, q0h));
        const Quat Cq = qmul(qs0, qconj(q0h));
        qmat(Cq, MC);
        double rx, ry, rz;
        mat_vec(R, pos[0], pos[1], pos[2], rx, ry, rz);
        hdr[0] = s[0] * rx + t[0]; hdr[1] = s[0] * ry + t[1]; hdr[2] = s[0] * rz + t[2];
        hdr[5] = Cq.x; hdr[6] = Cq.y; hdr[7] = Cq.z; hdr[8] = Cq.w;
        status[0] = st;
    }
    __syncthreads();
    const double t_eval = ts[0] + prm->eval_skip;
    // ordered compaction of the evaluation set (ts > ts[0] + skip; every pose is valid here), in slabs of 1024
    for (int lo = 0; lo < n; lo += 1024) {
        const int i = lo + threadIdx.x;
        const bool keep = i < n && ts[i] > t_eval;
        const unsigned bal = __ballot_sync(GSF_FULL_MASK, keep);
        __shared__ int wcnt[32];
        if ((threadIdx.x & 31) == 0) wcnt[threadIdx.x >> 5] = __popc(bal);
        __syncthreads();
        int base = s_count;
        for (int w = 0; w < (threadIdx.x >> 5); ++w) base += wcnt[w];
        const int p = base + __popc(bal & ((1u << (threadIdx.x & 31)) - 1));
        if (keep && p < cap_pow2) { key[p] = z[3 * i]; idx[p] = i; }
        __syncthreads();
        if (threadIdx.x == 0) { int tot = 0; for (int w = 0; w < 32; ++w) tot += wcnt[w]; s_count += tot; }
        __syncthreads();
    }
    const int m = min(s_count, cap_pow2);
    for (int k = m + threadIdx.x; k < cap_pow2; k += 1024) { key[k] = INFINITY; idx[k] = 0x7fffffff; }
    __syncthreads();
    // bitonic sort by (x, pose index): ties keep pose order -> deterministic
    for (int k = 2; k <= cap_pow2; k <<= 1)
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = threadIdx.x; i < cap_pow2; i += 1024) {
                const int l = i ^ j;
                if (l > i) {
                    const double a = key[i], c = key[l];
                    const int ia = idx[i], ic = idx[l];
                    const bool up = (i & k) == 0;
                    const bool gt = a > c || (a == c && ia > ic);
                    if (gt == up) { key[i] = c; key[l] = a; idx[i] = ic; idx[l] = ia; }
                }
            }
            __syncthreads();
        }
    // step records; rank defaults to -1
    for (int i = threadIdx.x; i < n; i += 1024) {
        double* r = rec + (size_t)GRID_REC * i;
        if (i == 0) { r[0] = 0.0; r[1] = r[2] = r[3] = 0.0; }
        else {
            r[0] = fmax(1e-6, ts[i] - ts[i - 1]);
            mat_vec(MC, pos[3 * i] - pos[3 * i - 3], pos[3 * i + 1] - pos[3 * i - 2], pos[3 * i + 2] - pos[3 * i - 1], r[1], r[2], r[3]);
        }
        r[4] = z[3 * i]; r[5] = z[3 * i + 1]; r[6] = z[3 * i + 2]; r[7] = -1.0;
    }
    __syncthreads();
    for (int k = threadIdx.x; k < m; k += 1024) {
        const int i = idx[k];
        cand[3 * k] = z[3 * i]; cand[3 * k + 1] = z[3 * i + 1]; cand[3 * k + 2] = z[3 * i + 2];
        rec[(size_t)GRID_REC * i + 7] = (double)k;
    }
    if (threadIdx.x == 0) { hdr[3] = (double)m; hdr[4] = (double)s_count; }
}

// ----------------------------------------------------------------------------- the grid kernel: one thread per hypothesis
template <int GRID_THREADS>
__global__ void __launch_bounds__(GRID_THREADS) ekf_grid_kernel(const double* __restrict__ rec, const double* __restrict__ cand,
                                                                const double* __restrict__ hdr, int n, const FuseParams* __restrict__ params,
                                                                int H, double* __restrict__ err, double* __restrict__ stats,
                                                                const int* __restrict__ status) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double* tile = reinterpret_cast<double*>(smem_raw);                 // 2 x GRID_TILE x GRID_REC
    uint64_t* mbar = reinterpret_cast<uint64_t*>(tile + 2 * GRID_TILE * GRID_REC);
    double* cs = reinterpret_cast<double*>(mbar + 2);                   // candidates [m,3]
    const int m = (int)hdr[3];
    const int h = blockIdx.x * blockDim.x + threadIdx.x;               // blockDim.x <= GRID_THREADS (see launch_grid_t)
    const bool live = h < H;
    if ((status[0] & GRID_FATAL) || (int)hdr[4] != m) {                  // prerequisites failed / candidate set too large
        if (live) { stats[4 * (size_t)h] = stats[4 * (size_t)h + 1] = stats[4 * (size_t)h + 2] = nan(""); stats[4 * (size_t)h + 3] = 0.0; }
        return;
    }
    if (threadIdx.x == 0) { mbar_init(mbar, 1); mbar_init(mbar + 1, 1); fence_mbar_init(); }
    for (int k = threadIdx.x; k < 3 * m; k += blockDim.x) cs[k] = cand[k];
    __syncthreads();
    const int ntiles = (n + GRID_TILE - 1) / GRID_TILE;
    auto issue = [&](int tl) {
        const int cnt = min(GRID_TILE, n - tl * GRID_TILE);
        const uint32_t bytes = (uint32_t)cnt * GRID_REC * 8u;
        mbar_expect_tx(mbar + (tl & 1), bytes);
        bulk_g2s(tile + (size_t)(tl & 1) * GRID_TILE * GRID_REC, rec + (size_t)tl * GRID_TILE * GRID_REC, bytes, mbar + (tl & 1));
    };
    if (threadIdx.x == 0) { issue(0); if (ntiles > 1) issue(1); }

    const FuseParams* pr = params + (live ? h : 0);
    const double p0x = pr->p0[0], p0y = pr->p0[1], p0z = pr->p0[2];
    const double qx = pr->q[0], qy = pr->q[1], qz = pr->q[2], rx = pr->r[0], ry = pr->r[1], rz = pr->r[2];
    double x0 = hdr[0], x1 = hdr[1], x2 = hdr[2];
    double Px = p0x, Py = p0y, Pz = p0z;
    double se = 0.0, se2 = 0.0;
    int ne = 0;
    for (int tl = 0; tl < ntiles; ++tl) {
        mbar_wait(mbar + (tl & 1), (uint32_t)(tl >> 1) & 1u);
        const double* T = tile + (size_t)(tl & 1) * GRID_TILE * GRID_REC;
        const int cnt = min(GRID_TILE, n - tl * GRID_TILE);
#pragma unroll 1
        for (int k = 0; k < cnt; ++k) {
            const double* r = T + GRID_REC * k;
            const int i = tl * GRID_TILE + k;
            if (i > 0) {
                const double dt = r[0];
                {   const double pp = Px + qx * dt, kk = pp * rcp_(pp + rx), om = 1.0 - kk;
                    Px = om * pp * om + kk * rx * kk; x0 = om * (x0 + r[1]) + kk * r[4]; }
                {   const double pp = Py + qy * dt, kk = pp * rcp_(pp + ry), om = 1.0 - kk;
                    Py = om * pp * om + kk * ry * kk; x1 = om * (x1 + r[2]) + kk * r[5]; }
                {   const double pp = Pz + qz * dt, kk = pp * rcp_(pp + rz), om = 1.0 - kk;
                    Pz = om * pp * om + kk * rz * kk; x2 = om * (x2 + r[3]) + kk * r[6]; }
            }
            const int rank = (int)r[7];
            if (rank >= 0) {
                // exact nearest neighbour: own measurement first, then outwards in x order while the 1-D gap can still win
                double dx = x0 - cs[3 * rank], dy = x1 - cs[3 * rank + 1], dz = x2 - cs[3 * rank + 2];
                double best = dist2_rn(dx, dy, dz);
                for (int c = rank - 1; c >= 0; --c) {
                    dx = x0 - cs[3 * c];
                    if (dx > 0.0 && dx * dx >= best) break;
                    dy = x1 - cs[3 * c + 1]; dz = x2 - cs[3 * c + 2];
                    best = fmin(best, dist2_rn(dx, dy, dz));
                }
                for (int c = rank + 1; c < m; ++c) {
                    dx = cs[3 * c] - x0;
                    if (dx > 0.0 && dx * dx >= best) break;
                    dy = x1 - cs[3 * c + 1]; dz = x2 - cs[3 * c + 2];
                    best = fmin(best, dist2_rn(dx, dy, dz));
                }
                const double e = sqrt(best);
                if (live) err[(size_t)ne * H + h] = e;
                se += e; se2 += e * e; ++ne;
            }
        }
        __syncthreads();                                                // every thread is done with this tile buffer
        if (threadIdx.x == 0 && tl + 2 < ntiles) issue(tl + 2);
    }
    if (live) {
        double* o = stats + 4 * (size_t)h;
        o[0] = ne ? se / ne : nan(""); o[2] = ne ? sqrt(se2 / ne) : nan(""); o[3] = (double)ne;
        if (!ne) o[1] = nan("");
    }
}

// ----------------------------------------------------------------------------- median of each hypothesis' error column
// Exact order statistics by radix selection on the bit patterns (errors are >= 0, so the IEEE patterns order like
// unsigned integers): 8-bit digits from the top, per pass a 256-bin histogram of the digit among the elements that
// still match the selected prefix, then the bin holding the wanted rank.  One block handles MED_COLS = 32 adjacent
// hypotheses -- a warp reads 256 contiguous bytes of a table row, eight rows per block and step (the first version
// took four hypotheses per block, one 32-byte sector per row: 1.9 TB/s) -- and both middle ranks (np.median of an
// even count averages them) at once.  Histograms: [rank][column][257] (padded: equal digits in the 32 columns of a
// warp fall into 32 different banks).  Passes whose digit is identical for the whole column (sign / exponent bytes,
// typically) are skipped through the column's min / max keys; the bin of a rank is found by a warp (8 bins per lane,
// shuffle scan) instead of a 255-step serial walk.
constexpr int MED_COLS = 32;
constexpr int MED_HSTRIDE = 257;
constexpr int MED_CAP = 16;           // candidates per (column, rank) below which the selection finishes by gathering them
constexpr size_t MED_SMEM = (size_t)2 * MED_COLS * MED_HSTRIDE * sizeof(unsigned int);
__global__ void __launch_bounds__(256) grid_median_kernel(const double* __restrict__ err, const double* __restrict__ hdr, int H,
                                                          double* __restrict__ stats, const int* __restrict__ status) {
    extern __shared__ __align__(128) unsigned char smem_raw[];          // (same declaration as the grid kernel's)
    unsigned int* const hist = reinterpret_cast<unsigned int*>(smem_raw);      // [(w * MED_COLS + c) * MED_HSTRIDE + digit]
    __shared__ unsigned long long prefix[MED_COLS][2], kmin[MED_COLS], kmax[MED_COLS];
    __shared__ unsigned int rank[MED_COLS][2], pop[MED_COLS][2], lcount[MED_COLS][2];
    __shared__ unsigned long long list[MED_COLS][2][MED_CAP];
    __shared__ int first_s;
    const int m = (int)hdr[3];
    if ((status[0] & GRID_FATAL) || (int)hdr[4] != m || m == 0) return;
    const int ntiles = (H + MED_COLS - 1) / MED_COLS;
    const int tid = threadIdx.x, c = tid & 31, rg = tid >> 5;           // column of the tile, row group (0-7)
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int h0 = tile * MED_COLS;
        const bool live = h0 + c < H;
        const double* __restrict__ col = err + h0 + c;
        __syncthreads();
        if (tid < MED_COLS) { kmin[tid] = ~0ull; kmax[tid] = 0ull; rank[tid][0] = (unsigned)((m - 1) / 2); rank[tid][1] = (unsigned)(m / 2); prefix[tid][0] = prefix[tid][1] = 0ull; }
        __syncthreads();
        // column min / max keys
        if (live) {
            unsigned long long lo = ~0ull, hi = 0ull;
            for (int r = rg; r < m; r += 8) {
                const unsigned long long k = (unsigned long long)__double_as_longlong(__ldg(col + (size_t)r * H));
                lo = min(lo, k); hi = max(hi, k);
            }
            atomicMin(&kmin[c], lo); atomicMax(&kmax[c], hi);
        }
        __syncthreads();
        // first digit position (from the top) where any column of the tile varies
        if (tid < 32) {
            const unsigned long long x = h0 + tid < H ? (kmin[tid] ^ kmax[tid]) : 0ull;
            int d = x ? (__clzll((long long)x) >> 3) : 8;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) d = min(d, __shfl_xor_sync(GSF_FULL_MASK, d, o));
            if (tid == 0) first_s = d;
        }
        __syncthreads();
        const int first = first_s;
        if (tid < MED_COLS) {                                             // bytes above `first` are common to the whole column
            const unsigned long long keep = first == 0 ? 0ull : (~0ull << (64 - 8 * first));
            prefix[tid][0] = prefix[tid][1] = kmin[tid] & keep;
        }
        __syncthreads();
        for (int d = first; d < 8; ++d) {
            const int shift = 56 - 8 * d;
            for (int k = tid; k < 2 * MED_COLS * MED_HSTRIDE; k += 256) hist[k] = 0u;
            __syncthreads();
            if (live) {
                const unsigned long long himask = d == 0 ? 0ull : (~0ull << (64 - 8 * d));
                const unsigned long long pf0 = prefix[c][0], pf1 = prefix[c][1];
                unsigned int* const h0p = hist + (size_t)c * MED_HSTRIDE;
                unsigned int* const h1p = hist + (size_t)(MED_COLS + c) * MED_HSTRIDE;
                for (int r = rg; r < m; r += 8) {
                    const unsigned long long k = (unsigned long long)__double_as_longlong(__ldg(col + (size_t)r * H));
                    const unsigned dig = (unsigned)(k >> shift) & 255u;
                    const unsigned long long top = k & himask;
                    if (top == pf0) atomicAdd(h0p + dig, 1u);
                    if (top == pf1) atomicAdd(h1p + dig, 1u);
                }
            }
            __syncthreads();
            // one warp per (column, rank) pair, 8 pairs per warp: 8 bins per lane, inclusive shuffle scan, then the lane
            // whose range holds the rank walks its 8 bins
            for (int pair = rg; pair < 2 * MED_COLS; pair += 8) {
                const int q = pair & (MED_COLS - 1), w = pair >> 5;
                const unsigned int* hp = hist + (size_t)(w * MED_COLS + q) * MED_HSTRIDE + 8 * c;      // c = lane
                unsigned cnt[8], tot = 0;
#pragma unroll
                for (int k = 0; k < 8; ++k) { cnt[k] = hp[k]; tot += cnt[k]; }
                const unsigned r = rank[q][w];                          // read by every lane before the shuffles below, written after them
                unsigned inc = tot;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) { const unsigned y = __shfl_up_sync(GSF_FULL_MASK, inc, o); if (c >= o) inc += y; }
                const unsigned exc = inc - tot;
                // the owning lane: exc <= r < inc; if the rank lies beyond every bin (cannot happen for r < population)
                // the last lane takes it, like the serial walk that stopped at bin 255
                const bool mine = (r >= exc && r < inc) || (c == 31 && r >= inc);
                if (mine) {
                    unsigned acc = exc; int bin = 0;
                    for (; bin < 7; ++bin) { if (acc + cnt[bin] > r) break; acc += cnt[bin]; }
                    rank[q][w] = r - acc;
                    pop[q][w] = cnt[bin];
                    prefix[q][w] |= (unsigned long long)(8 * c + bin) << shift;
                }
            }
            __syncthreads();
            // Few candidates left in every column of the tile: one more pass gathers them (instead of up to six more
            // histogram passes) and the wanted rank is picked among at most MED_CAP keys.
            bool few = true;
            if (tid < 2 * MED_COLS && h0 + (tid & 31) < H) few = pop[tid & 31][tid >> 5] <= (unsigned)MED_CAP;
            if (tid < 2 * MED_COLS) lcount[tid & 31][tid >> 5] = 0u;
            if (d < 7 && __syncthreads_and(few)) {
                if (live) {
                    const unsigned long long mask2 = ~0ull << shift;
                    const unsigned long long pf0 = prefix[c][0], pf1 = prefix[c][1];
                    for (int r = rg; r < m; r += 8) {
                        const unsigned long long k = (unsigned long long)__double_as_longlong(__ldg(col + (size_t)r * H));
                        const unsigned long long top = k & mask2;
                        if (top == pf0) { const unsigned sl = atomicAdd(&lcount[c][0], 1u); if (sl < (unsigned)MED_CAP) list[c][0][sl] = k; }
                        if (top == pf1) { const unsigned sl = atomicAdd(&lcount[c][1], 1u); if (sl < (unsigned)MED_CAP) list[c][1][sl] = k; }
                    }
                }
                __syncthreads();
                if (tid < 2 * MED_COLS && h0 + (tid & 31) < H) {
                    const int q = tid & 31, w = tid >> 5;
                    const unsigned nl = min(lcount[q][w], (unsigned)MED_CAP), r = rank[q][w];
                    for (unsigned i = 0; i < nl; ++i) {
                        const unsigned long long x = list[q][w][i];
                        unsigned less = 0;
                        for (unsigned j2 = 0; j2 < nl; ++j2) { const unsigned long long y = list[q][w][j2]; less += (y < x || (y == x && j2 < i)) ? 1u : 0u; }
                        if (less == r) { prefix[q][w] = x; break; }
                    }
                }
                __syncthreads();
                break;
            }
        }
        if (tid < MED_COLS && h0 + tid < H) {
            const int h = h0 + tid;
            double* o = stats + 4 * (size_t)h;
            const double a = __longlong_as_double((long long)prefix[tid][0]), b = __longlong_as_double((long long)prefix[tid][1]);
            o[1] = (o[0] != o[0]) ? nan("") : ((m & 1) ? a : 0.5 * (a + b));     // a NaN error (NaN mean) -> NaN median, like np.median
        }
    }
}

// ============================================================================= separable noise grid (config 5 proper)
// The hypotheses of BASELINE config 5 are a PRODUCT grid (q_xy, q_z, r) with P0 and everything else shared.  P0, Q and R
// are diagonal (EKFGPSSLAM.py:684-686) and H = [I3 0], so the x / y / z components of the filter are three independent
// scalar filters: the x and y tracks depend on (q_xy, r) only, the z track on (q_z, r) only.  Kq Kr x/y tracks and
// Kz Kr z tracks (2 x 64^2 + 64^2 = 12 288 scalar filter runs) replace the 3 x 64^3 = 786 432 the per-hypothesis
// kernel above executes -- same arithmetic per step, so every track equals the corresponding component of ekf_grid_kernel
// to rounding -- and a combine kernel scores every (q_xy, q_z, r) triple against the candidate set.
//   tracks_*_kernel      the scalar filter runs, parallel in time (see below); 32 x 32 tiles are transposed through shared
//                        memory so that the table [track][pose] is written in 256-byte rows;
//   grid_combine_kernel  one block per hypothesis at a time (threads over poses -- the tracks make the poses of a
//                        hypothesis independent, so the evaluation is parallel in time): coalesced reads of the three
//                        track rows (L2-resident by the loop order), exact pruned nearest neighbour in the x-sorted
//                        candidate set (shared memory), errors kept in shared memory, mean / RMSE by a fixed-order block
//                        reduction, median by radix selection in shared memory.  No [poses, hypotheses] error table.
static int pow2_at_least(long long n);
// Tracks, parallel in time.  A scalar filter run is a 4541-step recursion; one thread per track makes the launch one long
// dependent chain (0.79 ms, whatever the number of tracks -- the part of a rank's step that sharding over 8 GPUs does not
// shrink).  The recursion is two scans: the covariance map p -> r(p + q dt)/(p + q dt + r) is a Moebius map (2x2 step
// matrix [1 qa; g g qa + 1], g = 1/r), the state map x -> (1 - k)(x + u) + k z is affine.  Three launches over
// (track, chunk of TRK_CHUNK steps):
//   A  composite Moebius matrix of every chunk;
//   B  covariance at the chunk start (product of the chunks before it, applied to P0), then the exact per-step recursion of
//      ekf_grid_kernel through the chunk, composing the chunk's affine state map;
//   C  state at the chunk start (composition of the chunks before it, applied to x0), the recursion once more with the
//      state, 32 x 32 tiles transposed through shared memory into the [track][pose] table.
// Inside a chunk every step is the arithmetic of ekf_grid_kernel; only the chunk-start values come from the scans, and the
// covariance recursion is a contraction, so the tracks agree with the serial ones to a few 1e-16 relative.
constexpr int TRK_THREADS = 128;
constexpr int TRK_CHUNK = 128;        // steps per chunk (a multiple of 32)

struct TrackSet {
    const double* rec; const double* hdr; const FuseParams* base; const double* qxy; const double* qz; const double* rr;
    int n, Kr, iq0, nq, iz0, nz, nchunks;
    const int* status;
};
__device__ __forceinline__ void track_params(const TrackSet& T, int t, int& axis, double& q, double& r) {
    const int nxy = T.nq * T.Kr;
    if (t < 2 * nxy) { axis = t / nxy; const int rem = t - axis * nxy; q = T.qxy[T.iq0 + rem / T.Kr]; r = T.rr[rem % T.Kr]; }
    else { axis = 2; const int rem = t - 2 * nxy; q = T.qz[T.iz0 + rem / T.Kr]; r = T.rr[rem % T.Kr]; }
}
__device__ __forceinline__ void moeb2_rescale(double* m) {
    const int hi = max(max(__double2hiint(m[0]), __double2hiint(m[1])), max(__double2hiint(m[2]), __double2hiint(m[3])));
    const int e = ((hi >> 20) & 0x7ff) - 1023;
    const double sc = __hiloint2double((1023 - e) << 20, 0);
    m[0] *= sc; m[1] *= sc; m[2] *= sc; m[3] *= sc;
}
// thread (track, chunk); layout of the chunk arrays: [chunk][track] (coalesced over the tracks of a warp)
__global__ void __launch_bounds__(TRK_THREADS) tracks_moebius_kernel(const TrackSet T, int NT, double* __restrict__ mo) {
    if (T.status[0] & GRID_FATAL) return;
    const int t = blockIdx.x * TRK_THREADS + threadIdx.x, c = blockIdx.y;
    if (t >= NT) return;
    int axis; double q, r;
    track_params(T, t, axis, q, r);
    const double g = 1.0 / r;
    double m[4] = {1.0, 0.0, 0.0, 1.0};
    const int i0 = max(c * TRK_CHUNK, 1), i1 = min((c + 1) * TRK_CHUNK, T.n);
    for (int i = i0; i < i1; ++i) {
        const double qa = q * T.rec[(size_t)GRID_REC * i];
        m[0] = fma(qa, m[2], m[0]); m[1] = fma(qa, m[3], m[1]);       // [1 qa; 0 1] * m
        m[2] = fma(g, m[0], m[2]); m[3] = fma(g, m[1], m[3]);         // [1 0; g 1] * that
        if (((i - i0) & 15) == 15) moeb2_rescale(m);
    }
    moeb2_rescale(m);
    double* o = mo + ((size_t)c * NT + t) * 4;
    o[0] = m[0]; o[1] = m[1]; o[2] = m[2]; o[3] = m[3];
}
__device__ __forceinline__ double tracks_chunk_start_cov(const TrackSet& T, int NT, const double* __restrict__ mo, int t, int c, double P0) {
    double P = P0;
    for (int k = 0; k < c; ++k) {                                     // the chunks before this one, in order
        const double* m = mo + ((size_t)k * NT + t) * 4;
        P = (m[0] * P + m[1]) * rcp_(m[2] * P + m[3]);
    }
    return P;
}
__global__ void __launch_bounds__(TRK_THREADS) tracks_affine_kernel(const TrackSet T, int NT, const double* __restrict__ mo,
                                                                     double* __restrict__ pstart, double* __restrict__ aff) {
    if (T.status[0] & GRID_FATAL) return;
    const int t = blockIdx.x * TRK_THREADS + threadIdx.x, c = blockIdx.y;
    if (t >= NT) return;
    int axis; double q, r;
    track_params(T, t, axis, q, r);
    double P = tracks_chunk_start_cov(T, NT, mo, t, c, T.base->p0[axis]);
    pstart[(size_t)c * NT + t] = P;
    double A = 1.0, B = 0.0;
    const int i0 = max(c * TRK_CHUNK, 1), i1 = min((c + 1) * TRK_CHUNK, T.n);
    for (int i = i0; i < i1; ++i) {
        const double* rc = T.rec + (size_t)GRID_REC * i;
        const double pp = P + q * rc[0], kk = pp * rcp_(pp + r), om = 1.0 - kk;
        P = om * pp * om + kk * r * kk;
        B = om * (B + rc[1 + axis]) + kk * rc[4 + axis]; A *= om;
    }
    aff[((size_t)c * NT + t) * 2] = A; aff[((size_t)c * NT + t) * 2 + 1] = B;
}
__global__ void __launch_bounds__(TRK_THREADS) tracks_write_kernel(const TrackSet T, int NT, const double* __restrict__ pstart,
                                                                    const double* __restrict__ aff, double* __restrict__ tracks, long long npad) {
    __shared__ double tile[TRK_THREADS / 32][32][33];
    if (T.status[0] & GRID_FATAL) return;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int t = blockIdx.x * TRK_THREADS + threadIdx.x, c = blockIdx.y;
    const int tw0 = t - lane;
    if (tw0 >= NT) return;
    const int tt = t < NT ? t : NT - 1;
    int axis; double q, r;
    track_params(T, tt, axis, q, r);
    double x = T.hdr[axis];
    for (int k = 0; k < c; ++k) x = aff[((size_t)k * NT + tt) * 2] * x + aff[((size_t)k * NT + tt) * 2 + 1];
    double P = pstart[(size_t)c * NT + tt];
    const int c0 = c * TRK_CHUNK, c1 = min((c + 1) * TRK_CHUNK, T.n);
    for (int i0 = c0; i0 < c1; i0 += 32) {
        const int cnt = min(32, c1 - i0);
        for (int k = 0; k < cnt; ++k) {
            const int i = i0 + k;
            if (i > 0) {
                const double* rc = T.rec + (size_t)GRID_REC * i;
                const double pp = P + q * rc[0], kk = pp * rcp_(pp + r), om = 1.0 - kk;     // the step of ekf_grid_kernel, verbatim
                P = om * pp * om + kk * r * kk; x = om * (x + rc[1 + axis]) + kk * rc[4 + axis];
            }
            tile[warp][k][lane] = x;
        }
        __syncwarp();
        if (lane < cnt) {
            for (int tr = 0; tr < 32 && tw0 + tr < NT; ++tr) tracks[(size_t)(tw0 + tr) * npad + i0 + lane] = tile[warp][lane][tr];
        }
        __syncwarp();
    }
}

// Candidate order for the combine kernel: own[k] = pose of the k-th candidate (x order); hgap[2k], hgap[2k+1] = a quarter of
// the squared distance from candidate k to its nearest / third nearest other candidate; near2[2k..] = its two nearest.
// Triangle inequality: a query at distance e from its own measurement z_k can only be beaten by candidates
// within 2e of z_k -- none if e is under half the nearest distance, at most the two nearest if under half the third.
__global__ void __launch_bounds__(256) grid_prep_order_kernel(const double* __restrict__ rec, const double* __restrict__ cand,
                                                              const double* __restrict__ hdr, int n, int* __restrict__ own,
                                                              float* __restrict__ hgap, unsigned short* __restrict__ near2,
                                                              const int* __restrict__ status) {
    if (status[0] & GRID_FATAL) return;
    const int m = (int)hdr[3];
    if ((int)hdr[4] != m) return;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const int r = (int)rec[(size_t)GRID_REC * i + 7];
        if (r >= 0) own[r] = i;
    }
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < m; k += gridDim.x * blockDim.x) {
        // the three nearest other candidates of candidate k (pruned scan in x order)
        const double x0 = cand[3 * k], x1 = cand[3 * k + 1], x2 = cand[3 * k + 2];
        double b1 = INFINITY, b2 = INFINITY, b3 = INFINITY;
        int i1 = k, i2 = k;
        auto take = [&](int c) {
            const double d = dist2_rn(x0 - cand[3 * c], x1 - cand[3 * c + 1], x2 - cand[3 * c + 2]);
            if (d < b1) { b3 = b2; b2 = b1; i2 = i1; b1 = d; i1 = c; }
            else if (d < b2) { b3 = b2; b2 = d; i2 = c; }
            else if (d < b3) b3 = d;
        };
        for (int c = k - 1; c >= 0; --c) { const double dx = x0 - cand[3 * c]; if (dx * dx >= b3) break; take(c); }
        for (int c = k + 1; c < m; ++c) { const double dx = cand[3 * c] - x0; if (dx * dx >= b3) break; take(c); }
        // skip bounds, rounded down: a query within half the distance to the nearest other candidate keeps its own measurement;
        // within half the distance to the third nearest, only the two nearest need a look
        hgap[2 * k] = b1 < INFINITY ? __double2float_rd(0.25 * b1 * (1.0 - 1e-6)) : 0.0f;
        hgap[2 * k + 1] = b3 < INFINITY ? __double2float_rd(0.25 * b3 * (1.0 - 1e-6)) : 0.0f;
        near2[2 * k] = (unsigned short)i1; near2[2 * k + 1] = (unsigned short)i2;
    }
}
// Second scan order for the combine kernel.  The 1-D pruning of an x-sorted scan degenerates where the track runs along y
// (all the poses of that stretch share nearly the same x): such candidates are scanned in y order instead.  One block:
// bitonic sort of the candidates' (y, x-rank) pairs -> yperm[j] = x-rank of the j-th candidate in y order; yrank[k] = position of
// candidate k in that order if its local track direction (from the measurements 4 poses before / after) is closer to the
// y axis than to the x axis, 0xFFFF otherwise (scan in x order).
__global__ void __launch_bounds__(1024) grid_prep_ysort_kernel(const double* __restrict__ cand, const double* __restrict__ z,
                                                               const double* __restrict__ hdr, int n, const int* __restrict__ own,
                                                               unsigned short* __restrict__ yperm, unsigned short* __restrict__ yrank,
                                                               const int* __restrict__ status, int cap_pow2) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double* key = reinterpret_cast<double*>(smem_raw);
    int* idx = reinterpret_cast<int*>(key + cap_pow2);
    if (status[0] & GRID_FATAL) return;
    const int m = (int)hdr[3];
    if ((int)hdr[4] != m || m > 65535) return;
    for (int k = threadIdx.x; k < cap_pow2; k += 1024) { key[k] = k < m ? cand[3 * k + 1] : INFINITY; idx[k] = k < m ? k : 0x7fffffff; }
    __syncthreads();
    for (int k = 2; k <= cap_pow2; k <<= 1)
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = threadIdx.x; i < cap_pow2; i += 1024) {
                const int l = i ^ j;
                if (l > i) {
                    const double a = key[i], c = key[l];
                    const int ia = idx[i], ic = idx[l];
                    const bool up = (i & k) == 0;
                    const bool gt = a > c || (a == c && ia > ic);
                    if (gt == up) { key[i] = c; key[l] = a; idx[i] = ic; idx[l] = ia; }
                }
            }
            __syncthreads();
        }
    for (int j = threadIdx.x; j < m; j += 1024) {
        const int k = idx[j];
        yperm[j] = (unsigned short)k;
        const int i = own[k], ia = max(i - 4, 0), ib = min(i + 4, n - 1);
        const double dx = fabs(z[3 * ib] - z[3 * ia]), dy = fabs(z[3 * ib + 1] - z[3 * ia + 1]);
        yrank[k] = dy > dx ? (unsigned short)j : (unsigned short)0xFFFF;
    }
}
// One block per track: the row (time order) is re-ordered in place to candidate (x) order through shared memory, so that
// the combine kernel reads tracks, candidates and errors with the same index.
__global__ void __launch_bounds__(256) grid_permute_kernel(double* __restrict__ tracks, long long npad, int n, const int* __restrict__ own,
                                                           const double* __restrict__ hdr, const int* __restrict__ status) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double* row = reinterpret_cast<double*>(smem_raw);
    if (status[0] & GRID_FATAL) return;
    const int m = (int)hdr[3];
    if ((int)hdr[4] != m) return;
    double* g = tracks + (size_t)blockIdx.x * npad;
    for (int i = threadIdx.x; i < n; i += blockDim.x) row[i] = g[i];
    __syncthreads();
    for (int k = threadIdx.x; k < m; k += blockDim.x) g[k] = row[own[k]];
}

#ifdef GSF_GRID_CLK
__device__ unsigned long long g_grid_clk[8];
#endif
constexpr int CMB_THREADS = 1024;
constexpr int CMB_GROUP = 8;          // q_xy values per loop group: their 2 x 8 x Kr x/y rows stay L2-resident across the q_z loop
struct CombineArgs {
    const double* cand; const float* hgap; const unsigned short* near2; const double* hdr; const double* tracks; long long npad;
    const unsigned short* yperm; const unsigned short* yrank;
    int n, Kz, Kr, iq0, nq, iz0, nz;
    long long h_first, h_count;
    double* stats; const int* status; unsigned long long* slot_counter;
};
struct CmbShared {
    double red[2][32];
    unsigned long long redk[2][32];
    double fin[2];
    unsigned long long klo, khi;
    SelectShared<CMB_THREADS, 1> sel;
};
// Exact pruned nearest-neighbour scan around candidate k: outwards in x order (or in y order where the track runs along y)
// while the 1-D gap alone can still beat the best squared distance.
__device__ __forceinline__ double full_scan(const double* __restrict__ cs, const unsigned short* __restrict__ yp,
                                            const unsigned short* __restrict__ yr, int m, int k, double x0, double x1, double x2, double best) {
    double dx, dy, dz;
    const int jr = yr[k];
    if (jr == 0xFFFF) {
        for (int c = k - 1; c >= 0; --c) {
            dx = x0 - cs[3 * c];
            if (dx > 0.0 && dx * dx >= best) break;
            dy = x1 - cs[3 * c + 1]; dz = x2 - cs[3 * c + 2];
            best = fmin(best, dist2_rn(dx, dy, dz));
        }
        for (int c = k + 1; c < m; ++c) {
            dx = cs[3 * c] - x0;
            if (dx > 0.0 && dx * dx >= best) break;
            dy = x1 - cs[3 * c + 1]; dz = x2 - cs[3 * c + 2];
            best = fmin(best, dist2_rn(dx, dy, dz));
        }
    } else {
        for (int j = jr - 1; j >= 0; --j) {
            const int c = yp[j];
            dy = x1 - cs[3 * c + 1];
            if (dy > 0.0 && dy * dy >= best) break;
            dx = x0 - cs[3 * c]; dz = x2 - cs[3 * c + 2];
            best = fmin(best, dist2_rn(dx, dy, dz));
        }
        for (int j = jr + 1; j < m; ++j) {
            const int c = yp[j];
            dy = cs[3 * c + 1] - x1;
            if (dy > 0.0 && dy * dy >= best) break;
            dx = x0 - cs[3 * c]; dz = x2 - cs[3 * c + 2];
            best = fmin(best, dist2_rn(dx, dy, dz));
        }
    }
    return best;
}
__global__ void __launch_bounds__(CMB_THREADS, 1) grid_combine_kernel(const CombineArgs A) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int m = (int)A.hdr[3];
    __shared__ CmbShared S;
    double* const cs = reinterpret_cast<double*>(smem_raw);              // candidates [m,3], x-sorted
    double* const err = cs + 3 * (size_t)m;                              // errors [m], candidate order
    float* const hg = reinterpret_cast<float*>(err + m);                 // skip bounds [m,2]
    unsigned short* const nn = reinterpret_cast<unsigned short*>(hg + 2 * (size_t)m);      // two nearest other candidates [m,2]
    unsigned short* const yp = nn + 2 * (size_t)m + ((2 * m) & 1 ? 1 : 0);         // y order -> x-rank [m]
    unsigned short* const yr = yp + ((m + 3) & ~3);                      // x-rank -> position in y order, 0xFFFF: scan in x order
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const long long total_slots = (long long)((A.nq + CMB_GROUP - 1) / CMB_GROUP) * CMB_GROUP * A.Kz * A.Kr;
    const bool dead = (A.status[0] & GRID_FATAL) || (int)A.hdr[4] != m || m == 0;
    if (!dead) {
        for (int k = tid; k < 3 * m; k += CMB_THREADS) cs[k] = A.cand[k];
        for (int k = tid; k < m; k += CMB_THREADS) {
            hg[2 * k] = A.hgap[2 * k]; hg[2 * k + 1] = A.hgap[2 * k + 1]; nn[2 * k] = A.near2[2 * k]; nn[2 * k + 1] = A.near2[2 * k + 1];
            yp[k] = A.yperm[k]; yr[k] = A.yrank[k];
        }
    }
    __syncthreads();
    const int nxy = A.nq * A.Kr;
    // hypotheses differ a lot in cost (slow filters have large errors and need full scans): slots are handed out by a counter,
    // not by a static stride
    __shared__ long long s_slot;
    for (;;) {
        __syncthreads();
        if (tid == 0) s_slot = (long long)atomicAdd(A.slot_counter, 1ull);
        __syncthreads();
        const long long slot = s_slot;
        if (slot >= total_slots) break;
        // slot -> (q_xy, q_z, r): groups of CMB_GROUP q_xy values outermost, then q_z, then r, q_xy within the group innermost
        // (32-bit arithmetic: the launcher refuses grids beyond 2^31 slots)
        const unsigned per_group = (unsigned)(CMB_GROUP * A.Kz * A.Kr);
        const unsigned us = (unsigned)slot;
        const int g = (int)(us / per_group);
        const unsigned sg = us - (unsigned)g * per_group;
        const int iql = g * CMB_GROUP + (int)(sg % CMB_GROUP);           // q_xy index relative to iq0
        const int ir = (int)((sg / CMB_GROUP) % (unsigned)A.Kr), iz = (int)(sg / (unsigned)(CMB_GROUP * A.Kr));
        if (iql >= A.nq) continue;
        const long long h = ((long long)(A.iq0 + iql) * A.Kz + iz) * A.Kr + ir;
        if (h < A.h_first || h >= A.h_first + A.h_count || iz < A.iz0 || iz >= A.iz0 + A.nz) continue;
        double* const o = A.stats + 4 * (size_t)(h - A.h_first);
        if (dead) { if (tid == 0) { o[0] = o[1] = o[2] = nan(""); o[3] = 0.0; } continue; }
        const double* __restrict__ tx = A.tracks + (size_t)(iql * A.Kr + ir) * A.npad;
        const double* __restrict__ ty = A.tracks + (size_t)(nxy + iql * A.Kr + ir) * A.npad;
        const double* __restrict__ tz = A.tracks + (size_t)(2 * nxy + (iz - A.iz0) * A.Kr + ir) * A.npad;
#ifdef GSF_GRID_CLK
        long long c0 = clock64();
#endif
        // ---- phase 1: nearest-neighbour error of every evaluated pose (:1028-1031), exact and pruned; queries in candidate order.
        // Three tiers by the triangle inequality: own measurement only / the two nearest of z_k / a full pruned scan.
        // (Queueing the rare full scans for a second, dense pass -- a single one stalls its warp -- measured slower than
        // leaving them in place: 26.0 vs 23.4 ms; the extra sweeps over err[] cost more than the divergence.)
        double se = 0.0, se2 = 0.0;
        unsigned long long klo = ~0ull, khi = 0ull;
        {
            int k = tid;
            double x0 = 0.0, x1 = 0.0, x2 = 0.0;
            if (k < m) { x0 = __ldg(tx + k); x1 = __ldg(ty + k); x2 = __ldg(tz + k); }
            while (k < m) {
                const int kn = k + CMB_THREADS;
                double n0 = 0.0, n1 = 0.0, n2 = 0.0;
                if (kn < m) { n0 = __ldg(tx + kn); n1 = __ldg(ty + kn); n2 = __ldg(tz + kn); }      // next query's track values (L2 latency)
                double best = dist2_rn(x0 - cs[3 * k], x1 - cs[3 * k + 1], x2 - cs[3 * k + 2]);
                if (!(best < (double)hg[2 * k])) {                       // another candidate may be nearer than the own measurement
                    if (best < (double)hg[2 * k + 1]) {                  // ... but only one of the two nearest of z_k
                        const int c1 = nn[2 * k], c2 = nn[2 * k + 1];
                        best = fmin(best, dist2_rn(x0 - cs[3 * c1], x1 - cs[3 * c1 + 1], x2 - cs[3 * c1 + 2]));
                        best = fmin(best, dist2_rn(x0 - cs[3 * c2], x1 - cs[3 * c2 + 1], x2 - cs[3 * c2 + 2]));
                    } else best = full_scan(cs, yp, yr, m, k, x0, x1, x2, best);
                }
                const double e = sqrt(best);
                err[k] = e;
                se += e; se2 += e * e;
                const unsigned long long key = (unsigned long long)__double_as_longlong(e);      // e >= 0 or NaN: patterns order like the values
                klo = min(klo, key); khi = max(khi, key);
                k = kn; x0 = n0; x1 = n1; x2 = n2;
            }
        }
        // fixed-order block reduction (lane tree, then warp tree)
#pragma unroll
        for (int ofs = 16; ofs > 0; ofs >>= 1) {
            se += __shfl_xor_sync(GSF_FULL_MASK, se, ofs); se2 += __shfl_xor_sync(GSF_FULL_MASK, se2, ofs);
            klo = min(klo, __shfl_xor_sync(GSF_FULL_MASK, klo, ofs)); khi = max(khi, __shfl_xor_sync(GSF_FULL_MASK, khi, ofs));
        }
        if (lane == 0) { S.red[0][warp] = se; S.red[1][warp] = se2; S.redk[0][warp] = klo; S.redk[1][warp] = khi; }
        __syncthreads();
        if (warp == 0) {
            double a = S.red[0][lane], b = S.red[1][lane];
            unsigned long long lo = S.redk[0][lane], hi = S.redk[1][lane];
#pragma unroll
            for (int ofs = 16; ofs > 0; ofs >>= 1) {
                a += __shfl_xor_sync(GSF_FULL_MASK, a, ofs); b += __shfl_xor_sync(GSF_FULL_MASK, b, ofs);
                lo = min(lo, __shfl_xor_sync(GSF_FULL_MASK, lo, ofs)); hi = max(hi, __shfl_xor_sync(GSF_FULL_MASK, hi, ofs));
            }
            if (lane == 0) { S.fin[0] = a; S.fin[1] = b; S.klo = lo; S.khi = hi; }
        }
        __syncthreads();
#ifdef GSF_GRID_CLK
        long long c2 = clock64();
#endif
        // ---- phase 2: exact median (np.median: mean of the two middle order statistics for an even count); the median of any
        //      sample lies within one standard deviation of its mean, which brackets the first level of the selection
        const bool has_nan = S.fin[0] != S.fin[0];
        double med = nan("");
        if (!has_nan) {
            const double mu = S.fin[0] / m, var = S.fin[1] / m - mu * mu, sg = var > 0.0 ? sqrt(var) : 0.0;
            block_select<CMB_THREADS, 1>(S.sel, err, m, (unsigned)((m - 1) / 2), !(m & 1), S.klo, S.khi, mu - sg, mu + sg);
            med = (m & 1) ? S.sel.v0 : 0.5 * (S.sel.v0 + S.sel.v1);
        }
        if (tid == 0) { o[0] = S.fin[0] / m; o[1] = med; o[2] = sqrt(S.fin[1] / m); o[3] = (double)m; }
        __syncthreads();
#ifdef GSF_GRID_CLK
        if (blockIdx.x == 0 && (tid & 31) == 0) {       // own phase-1 time per warp, wait at the barrier, median (warp 0)
            const long long c3 = clock64();
            atomicAdd(&g_grid_clk[0], (unsigned long long)(c1 - c0));
            atomicAdd(&g_grid_clk[1], (unsigned long long)(c2 - c1));
            if (tid == 0) { atomicAdd(&g_grid_clk[2], (unsigned long long)(c3 - c2)); atomicAdd(&g_grid_clk[3], (unsigned long long)(c3 - c0)); atomicAdd(&g_grid_clk[4], 1ull); }
            atomicMax(&g_grid_clk[5], (unsigned long long)(c1 - c0));
        }
#endif
    }
}

// work layout (doubles): the header block of the per-hypothesis path (hdr, R t s, offsets2, status, rec, cand, Umeyama tiles,
// mask), own (n ints), hgap (n), then the track table [tracks][npad].
static void noise_grid_ranges(int Kz, int Kr, long long h_first, long long h_count, int& iq0, int& nq, int& iz0, int& nz) {
    const long long slab = (long long)Kz * Kr;
    iq0 = (int)(h_first / slab);
    const int iq1 = (int)((h_first + h_count - 1) / slab);
    nq = iq1 - iq0 + 1;
    if (nq == 1) { iz0 = (int)((h_first % slab) / Kr); nz = (int)(((h_first + h_count - 1) % slab) / Kr) - iz0 + 1; }
    else { iz0 = 0; nz = Kz; }
}
static long long noise_grid_head_doubles(long long n) {
    return 16 + 16 + 2 + 2 + 8 * n + 3 * n + (long long)sim3_tiles_for(n) * 20 + (n + 7) / 8 + 4 + (n + 1) / 2 + n + (n + 1) / 2 + 2 * ((n + 3) / 4) + 8;
}
long long noise_grid_work_doubles(long long n, int Kq, int Kz, int Kr, long long h_first, long long h_count) {
    (void)Kq;
    int iq0, nq, iz0, nz;
    noise_grid_ranges(Kz, Kr, h_first, h_count, iq0, nq, iz0, nz);
    const long long npad = (n + 31) / 32 * 32;
    const long long NT = (2ll * nq + nz) * Kr, nchunks = (n + 127) / 128;
    return noise_grid_head_doubles(n) + NT * npad + 4 + 7 * nchunks * NT + 8;        // table + chunk scratch (Moebius 4, covariance 1, affine 2)
}
cudaError_t launch_noise_grid(const double* ts, const double* pos, const double* quat, const double* z, long long n,
                              const FuseParams* base, const double* qxy, const double* qz, const double* rr, int Kq, int Kz, int Kr,
                              long long h_first, long long h_count, double* work, double* stats, double* sim3_out, int* status_out,
                              int max_smem, int num_sms, cudaStream_t stream) {
    (void)Kq;
    double* hdr = work;
    double* R = work + 16; double* t = R + 9; double* s = t + 3;
    long long* offsets2 = reinterpret_cast<long long*>(work + 32);
    int* st = reinterpret_cast<int*>(work + 34);
    double* rec = work + 36;
    double* cand = rec + 8 * n;
    double* uwork = cand + 3 * n;
    unsigned char* mask = reinterpret_cast<unsigned char*>(uwork + (size_t)sim3_tiles_for(n) * 20);
    double* after = reinterpret_cast<double*>(mask) + (n + 7) / 8 + 4;
    int* own = reinterpret_cast<int*>(after);
    float* hgap = reinterpret_cast<float*>(after + (n + 1) / 2);                   // [n,2] floats = n doubles
    unsigned short* near2 = reinterpret_cast<unsigned short*>(after + (n + 1) / 2 + n);      // [n,2] = (n + 1) / 2 doubles
    unsigned short* yperm = near2 + 4 * ((n + 1) / 2);
    unsigned short* yrank = yperm + 4 * ((n + 3) / 4);
    double* tracks = work + noise_grid_head_doubles(n);
    tracks = reinterpret_cast<double*>((reinterpret_cast<uintptr_t>(tracks) + 31) & ~(uintptr_t)31);
    const long long npad = (n + 31) / 32 * 32;
    int iq0, nq, iz0, nz;
    noise_grid_ranges(Kz, Kr, h_first, h_count, iq0, nq, iz0, nz);
    double* tracks_scratch = tracks + (size_t)(2 * nq + nz) * Kr * npad + 4;
    const int cap2 = pow2_at_least(n);
    const size_t smem_prep = (size_t)cap2 * 12;
    const size_t smem_cmb = (size_t)n * 48 + 64;                           // candidates 24 + errors 8 + skip bounds 8 + two nearest 4 + y order 4 per pose
    if (smem_prep > (size_t)max_smem || smem_cmb > (size_t)max_smem || (size_t)n * 8 > (size_t)max_smem || n > 65535 ||
        (long long)((nq + CMB_GROUP - 1) / CMB_GROUP) * CMB_GROUP * Kz * Kr > 0x7fffffffll) return cudaErrorInvalidValue;
    grid_prep_select_kernel<<<1, 1024, 0, stream>>>(ts, z, (int)n, base, mask, offsets2, st);
    cudaError_t e = launch_umeyama(pos, z, offsets2, mask, 1, n, uwork, R, t, s, st + 1, stream);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(grid_prep_records_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_prep);
    if (e != cudaSuccess) return e;
    grid_prep_records_kernel<<<1, 1024, smem_prep, stream>>>(ts, pos, quat, z, (int)n, base, R, t, s, rec, cand, hdr, st, cap2);
    grid_prep_order_kernel<<<(int)((n + 255) / 256), 256, 0, stream>>>(rec, cand, hdr, (int)n, own, hgap, near2, st);
    e = cudaFuncSetAttribute(grid_prep_ysort_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_prep);
    if (e != cudaSuccess) return e;
    grid_prep_ysort_kernel<<<1, 1024, smem_prep, stream>>>(cand, z, hdr, (int)n, own, yperm, yrank, st, cap2);
    const int NT = (2 * nq + nz) * Kr;
    {
        TrackSet T;
        T.rec = rec; T.hdr = hdr; T.base = base; T.qxy = qxy; T.qz = qz; T.rr = rr; T.n = (int)n; T.Kr = Kr; T.iq0 = iq0; T.nq = nq; T.iz0 = iz0; T.nz = nz;
        T.nchunks = (int)((n + TRK_CHUNK - 1) / TRK_CHUNK); T.status = st;
        double* mo = tracks_scratch;                                  // [nchunks][NT][4]
        double* pstart = mo + (size_t)T.nchunks * NT * 4;             // [nchunks][NT]
        double* aff = pstart + (size_t)T.nchunks * NT;                // [nchunks][NT][2]
        const dim3 tg((NT + TRK_THREADS - 1) / TRK_THREADS, T.nchunks);
        tracks_moebius_kernel<<<tg, TRK_THREADS, 0, stream>>>(T, NT, mo);
        tracks_affine_kernel<<<tg, TRK_THREADS, 0, stream>>>(T, NT, mo, pstart, aff);
        tracks_write_kernel<<<tg, TRK_THREADS, 0, stream>>>(T, NT, pstart, aff, tracks, npad);
    }
    e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(grid_permute_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(n * 8));
    if (e != cudaSuccess) return e;
    grid_permute_kernel<<<NT, 256, (size_t)n * 8, stream>>>(tracks, npad, (int)n, own, hdr, st);
    CombineArgs ca;
    ca.cand = cand; ca.hgap = hgap; ca.near2 = near2; ca.hdr = hdr; ca.tracks = tracks; ca.npad = npad; ca.yperm = yperm; ca.yrank = yrank;
    ca.n = (int)n; ca.Kz = Kz; ca.Kr = Kr; ca.iq0 = iq0; ca.nq = nq; ca.iz0 = iz0; ca.nz = nz;
    ca.h_first = h_first; ca.h_count = h_count; ca.stats = stats; ca.status = st;
    ca.slot_counter = reinterpret_cast<unsigned long long*>(hdr + 12);           // hdr[12]: free header slot
    e = cudaMemsetAsync(ca.slot_counter, 0, sizeof(unsigned long long), stream);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(grid_combine_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_cmb);
    if (e != cudaSuccess) return e;
    long long blocks = h_count < num_sms ? h_count : num_sms;
    grid_combine_kernel<<<(unsigned)blocks, CMB_THREADS, smem_cmb, stream>>>(ca);
    e = cudaGetLastError();
    if (e != cudaSuccess) return e;
#ifdef GSF_GRID_CLK
    {
        unsigned long long hc[8];
        cudaStreamSynchronize(stream);
        cudaMemcpyFromSymbol(hc, g_grid_clk, sizeof(hc));
        const double nh = (double)(hc[4] ? hc[4] : 1);
        fprintf(stderr, "[grid clk] block 0, %llu hypotheses: per hypothesis -- phase 1 own time, mean over warps %.0f; wait+reduce %.0f; median %.0f; total %.0f cycles; longest phase 1 of any warp %llu\n",
                hc[4], hc[0] / nh / 32.0, hc[1] / nh / 32.0, hc[2] / nh, hc[3] / nh, hc[5]);
        unsigned long long zero[8] = {0};
        cudaMemcpyToSymbol(g_grid_clk, zero, sizeof(zero));
    }
#endif
    if (sim3_out) { e = cudaMemcpyAsync(sim3_out, R, 13 * sizeof(double), cudaMemcpyDeviceToDevice, stream); if (e != cudaSuccess) return e; }
    if (status_out) e = cudaMemcpyAsync(status_out, st, sizeof(int), cudaMemcpyDeviceToDevice, stream);
    return e;
}

// One block per SM (the candidate set fills most of the shared memory) and a block takes about the same time whatever
// its size (the per-hypothesis recursion is latency-bound): use the fewest waves that blocks of <= 1024 threads allow
// and size the blocks so that the waves are full -- 262 144 hypotheses on 148 SMs run as 2 x 148 blocks of 896
// threads (not 256 blocks of 1024: 1.73 waves that cost 2), 131 072 as one wave of 148 x 896.
static int grid_block_threads(int H, int num_sms) {
    if (num_sms <= 0) return 256;
    const long long waves = ((long long)H + (long long)num_sms * 1024 - 1) / ((long long)num_sms * 1024);
    const long long per_block = ((long long)H + waves * num_sms - 1) / (waves * num_sms);
    long long threads = (per_block + 31) / 32 * 32;
    if (threads > 1024) threads = 1024;
    if (threads < 32) threads = 32;
    return (int)threads;
}
template <int THREADS>
static cudaError_t launch_grid_t(const double* rec, const double* cand, const double* hdr, int n, const FuseParams* params, int H,
                                 double* err, double* stats, const int* st, size_t smem, int threads, cudaStream_t stream) {
    cudaError_t e = cudaFuncSetAttribute(ekf_grid_kernel<THREADS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    ekf_grid_kernel<THREADS><<<(H + threads - 1) / threads, threads, smem, stream>>>(rec, cand, hdr, n, params, H, err, stats, st);
    return cudaGetLastError();
}
static int pow2_at_least(long long n) { int p = 1; while (p < n) p <<= 1; return p; }

// work layout (doubles): hdr[16] | R[9] t[3] s[1] pad[3] | offsets2 (2 long long) | status (4 ints = 2 doubles) |
//                        rec[8 n] | cand[3 n] | umeyama tiles | mask (n bytes, padded) | err [n x H]
long long grid_work_doubles(long long n, int H) {
    return 16 + 16 + 2 + 2 + 8 * n + 3 * n + (long long)sim3_tiles_for(n) * 20 + (n + 7) / 8 + 4 + n * (long long)H;
}
cudaError_t launch_hypothesis_grid(const double* ts, const double* pos, const double* quat, const double* z, long long n,
                                   const FuseParams* params, int H, double* work, double* stats, double* sim3_out, int* status_out,
                                   int max_smem, int num_sms, cudaStream_t stream) {
    double* hdr = work;
    double* R = work + 16; double* t = R + 9; double* s = t + 3;
    long long* offsets2 = reinterpret_cast<long long*>(work + 32);
    int* st = reinterpret_cast<int*>(work + 34);
    double* rec = work + 36;
    double* cand = rec + 8 * n;
    double* uwork = cand + 3 * n;
    unsigned char* mask = reinterpret_cast<unsigned char*>(uwork + (size_t)sim3_tiles_for(n) * 20);
    double* err = reinterpret_cast<double*>(mask) + (n + 7) / 8;
    err = reinterpret_cast<double*>((reinterpret_cast<uintptr_t>(err) + 31) & ~(uintptr_t)31);      // 32-byte rows for the median kernel
    const int cap2 = pow2_at_least(n);
    const size_t smem_prep = (size_t)cap2 * 12;
    const size_t smem_grid = (size_t)2 * GRID_TILE * GRID_REC * 8 + 16 + (size_t)n * 24;
    if (smem_prep > (size_t)max_smem || smem_grid > (size_t)max_smem || false) return cudaErrorInvalidValue;
    grid_prep_select_kernel<<<1, 1024, 0, stream>>>(ts, z, (int)n, params, mask, offsets2, st);
    cudaError_t e = launch_umeyama(pos, z, offsets2, mask, 1, n, uwork, R, t, s, st + 1, stream);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(grid_prep_records_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_prep);
    if (e != cudaSuccess) return e;
    grid_prep_records_kernel<<<1, 1024, smem_prep, stream>>>(ts, pos, quat, z, (int)n, params, R, t, s, rec, cand, hdr, st, cap2);
    const int gthreads = grid_block_threads(H, num_sms);                   // the template bounds the registers per thread
    if (gthreads > 512) e = launch_grid_t<1024>(rec, cand, hdr, (int)n, params, H, err, stats, st, smem_grid, gthreads, stream);
    else if (gthreads > 256) e = launch_grid_t<512>(rec, cand, hdr, (int)n, params, H, err, stats, st, smem_grid, gthreads, stream);
    else e = launch_grid_t<256>(rec, cand, hdr, (int)n, params, H, err, stats, st, smem_grid, gthreads, stream);
    if (e != cudaSuccess) return e;
    const int ntiles = (H + MED_COLS - 1) / MED_COLS;
    const int mg = ntiles < num_sms * 3 ? ntiles : num_sms * 3;           // 66 KB of histograms per block: three blocks per SM
    e = cudaFuncSetAttribute(grid_median_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)MED_SMEM);
    if (e != cudaSuccess) return e;
    grid_median_kernel<<<mg, 256, MED_SMEM, stream>>>(err, hdr, H, stats, st);
    e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    if (sim3_out) { e = cudaMemcpyAsync(sim3_out, R, 13 * sizeof(double), cudaMemcpyDeviceToDevice, stream); if (e != cudaSuccess) return e; }
    if (status_out) e = cudaMemcpyAsync(status_out, st, sizeof(int), cudaMemcpyDeviceToDevice, stream);
    return e;
}

}  // namespace gsf
