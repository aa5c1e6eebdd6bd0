// K4: nearest-neighbour error statistics (evaluation step of /root/reference/EKFGPSSLAM.py:1013-1039).
//   idx   = {i : cand[i] has no NaN and ts[i] > ts[0] + skip}                       (:1015-1023)
//   e_i   = min_j |traj[idx_i] - cand[idx_j]|   (scipy cdist + np.min, :1030-1031)
//   stats = mean, median, sqrt(mean(e^2)), count                                    (:1033)
// The reference builds the full |idx| x |idx| distance matrix.  Here the minimum is found exactly but pruned:
// the candidates of a trajectory are bucketed along the longer horizontal axis of their bounding box (1024
// bins, counting sort with shared-memory atomics; the evaluation set is a SET, so the order inside a bin is
// free), and a query scans the contiguous run of bins within its own-measurement distance along that axis --
// the own measurement is always a candidate, so nothing farther can be the nearest.  The bounding box comes from every 8th
// pose (one input sweep less); the size of the evaluation set from the histogram.  Sums run in pose order per thread and are combined by a fixed tree (bit-reproducible); the
// median is an exact radix selection on the IEEE bit patterns (errors are >= 0, so they order like unsigned
// integers), with the digits above the first one that varies skipped.
// One block per trajectory at a time.  Candidates + errors live in shared memory (32 B per evaluation pose);
// trajectories beyond that capacity (~6900 poses) use the caller's workspace in global memory instead, so
// any length works -- the reference has no length limit either.
#include "gsf_common.cuh"
#include "gsf_internal.cuh"
#include "gsf_ptx.cuh"
#include "gsf_select.cuh"

namespace gsf {

#ifndef GSF_ATE_UNROLL
#define GSF_ATE_UNROLL 1                 // unroll factor of the input sweeps (tuning)
#endif
#ifndef GSF_ATE_BBOX_STRIDE
#define GSF_ATE_BBOX_STRIDE 8            // bounding box from every k-th pose (the bins clamp: exactness does not depend on it)
#endif
#ifndef GSF_ATE_QUNROLL
#define GSF_ATE_QUNROLL 1                // unroll factor of the query sweep
#endif
#ifndef GSF_ATE_MINB
#define GSF_ATE_MINB 4                   // 64 registers: four blocks per SM at 1000 poses (17.5 -> 10.7 ms on 262 144 trajectories)
#endif
constexpr int ATE_UNROLL = GSF_ATE_UNROLL;
constexpr int ATE_QUNROLL = GSF_ATE_QUNROLL;
constexpr int ATE_T = 256;               // threads per block
constexpr int ATE_NW = ATE_T / 32;
constexpr int ATE_NB = 1024;             // bins along the dominant axis

// fixed shared memory: bins (start offsets, then reused as cursors), histogram of the radix selection, scalars
struct AteShared {
    int bin[ATE_NB + 1];
    union {                              // the scatter cursors are dead when the selection starts
        int cursor[ATE_NB];
        SelectShared<ATE_T, 4> sel;
    };
    double red[4 * ATE_NW];
    unsigned long long kmin, kmax, above;
    unsigned int rank, le_count;
    unsigned long long prefix;
    int count, slot, any_nan, pad;
};

__device__ __forceinline__ unsigned long long ate_key(double e) { return (unsigned long long)__double_as_longlong(e); }

template <bool GLOBAL_WORK>
__global__ void __launch_bounds__(ATE_T, GSF_ATE_MINB) ate_nn_kernel(const AteArgs A) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    AteShared& S = *reinterpret_cast<AteShared*>(smem_raw);
    double* const smem_arr = reinterpret_cast<double*>(smem_raw + ((sizeof(AteShared) + 15) & ~(size_t)15));
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    for (int b = blockIdx.x; b < A.B; b += gridDim.x) {
        const long long e0 = A.offsets[b];
        const int n = (int)(A.offsets[b + 1] - e0);
        double* const o = A.stats + 4 * (size_t)b;
        const double* __restrict__ gc = A.cand + 3 * e0;
        const double* __restrict__ gt = A.traj + 3 * e0;
        const double* __restrict__ gts = A.ts + e0;
        double* const cs = GLOBAL_WORK ? A.work + 4 * e0 : smem_arr;          // bucketed candidates [m,3]
        double* const err = GLOBAL_WORK ? A.work + 4 * e0 + 3 * (size_t)n : smem_arr + 3 * (size_t)A.cap;   // errors [m]
        __syncthreads();
        // the next trajectory of this block into L2 now: its four sweeps (bounding box, histogram, scatter, queries) then read
        // L2 instead of waiting for HBM with a handful of loads in flight per thread
        if (tid == 0 && b + (int)gridDim.x < A.B) {
            const long long f0 = A.offsets[b + gridDim.x] & ~1ll, f1 = A.offsets[b + gridDim.x + 1] & ~1ll;      // even pose index: 16-byte aligned rows
            if (f1 > f0 && f1 - f0 < (1ll << 24)) {
                bulk_prefetch_l2(A.cand + 3 * f0, (uint32_t)((f1 - f0) * 24)); bulk_prefetch_l2(A.traj + 3 * f0, (uint32_t)((f1 - f0) * 24));
                bulk_prefetch_l2(A.ts + f0, (uint32_t)((f1 - f0) * 8));
            }
        }
        if (n <= 0) { if (tid == 0) { o[0] = o[1] = o[2] = nan(""); o[3] = 0.0; } continue; }
        const double t0 = gts[0] + A.skip;

        // ---- pass 1: size of the evaluation set and its bounding box in x / y
        for (int k = tid; k < ATE_NB + 1; k += ATE_T) S.bin[k] = 0;
        if (tid == 0) { S.slot = 0; S.any_nan = 0; S.kmin = ~0ull; S.kmax = 0ull; S.above = ~0ull; S.le_count = 0; }
        double lo_x = INFINITY, hi_x = -INFINITY, lo_y = INFINITY, hi_y = -INFINITY;
        int cnt = 0;
#pragma unroll ATE_UNROLL
        for (int i = tid * GSF_ATE_BBOX_STRIDE; i < n; i += ATE_T * GSF_ATE_BBOX_STRIDE) {
            const double cx = gc[3 * (size_t)i], cy = gc[3 * (size_t)i + 1], cz = gc[3 * (size_t)i + 2];
            if (!row_has_nan(cx, cy, cz) && gts[i] > t0) {
                ++cnt; lo_x = fmin(lo_x, cx); hi_x = fmax(hi_x, cx); lo_y = fmin(lo_y, cy); hi_y = fmax(hi_y, cy);
            }
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            cnt += __shfl_xor_sync(GSF_FULL_MASK, cnt, off);
            lo_x = fmin(lo_x, __shfl_xor_sync(GSF_FULL_MASK, lo_x, off)); hi_x = fmax(hi_x, __shfl_xor_sync(GSF_FULL_MASK, hi_x, off));
            lo_y = fmin(lo_y, __shfl_xor_sync(GSF_FULL_MASK, lo_y, off)); hi_y = fmax(hi_y, __shfl_xor_sync(GSF_FULL_MASK, hi_y, off));
        }
        if (lane == 0) { S.red[4 * warp] = lo_x; S.red[4 * warp + 1] = hi_x; S.red[4 * warp + 2] = lo_y; S.red[4 * warp + 3] = hi_y; S.cursor[warp] = cnt; }
        __syncthreads();
        int m = 0;
        lo_x = INFINITY; hi_x = -INFINITY; lo_y = INFINITY; hi_y = -INFINITY;
#pragma unroll
        for (int w = 0; w < ATE_NW; ++w) {
            m += S.cursor[w];
            lo_x = fmin(lo_x, S.red[4 * w]); hi_x = fmax(hi_x, S.red[4 * w + 1]); lo_y = fmin(lo_y, S.red[4 * w + 2]); hi_y = fmax(hi_y, S.red[4 * w + 3]);
        }
        __syncthreads();
        if (GSF_ATE_BBOX_STRIDE == 1 && (m == 0 || (!GLOBAL_WORK && m > A.cap))) {
            // m > cap cannot happen when the launcher sized cap from max_len >= n; reported, never silently wrong
            if (tid == 0) { o[0] = o[1] = o[2] = nan(""); o[3] = m > 0 ? -(double)m : 0.0; }
            continue;
        }
        if (GSF_ATE_BBOX_STRIDE > 1 && !(hi_x >= lo_x)) { lo_x = hi_x = lo_y = hi_y = 0.0; }      // no valid pose in the sample: one bin
        // dominant horizontal axis; bin k covers [base + k w, base + (k + 1) w)
        const int ax = (hi_y - lo_y > hi_x - lo_x) ? 1 : 0;
        const double base = ax ? lo_y : lo_x, ext = ax ? hi_y - lo_y : hi_x - lo_x;
        const double scale = ext > 0.0 ? (double)ATE_NB / ext : 0.0, wbin = ext / (double)ATE_NB;
        // rounding slack for the pruning bound: bin edges and the bin index of a candidate are each good to a few ulps of
        // the coordinate magnitude, so a bound is only trusted beyond that distance
        const double slack = 8.0 * 2.220446049250313e-16 * fmax(fabs(base), fabs(base + ext)) + 1e-300;

        // ---- pass 2: histogram, exclusive scan (warp 0 ... per thread 4 bins), pass 3: scatter
#pragma unroll ATE_UNROLL
        for (int i = tid; i < n; i += ATE_T) {
            const double cx = gc[3 * (size_t)i], cy = gc[3 * (size_t)i + 1], cz = gc[3 * (size_t)i + 2];
            if (!row_has_nan(cx, cy, cz) && gts[i] > t0) {
                int k = (int)(((ax ? cy : cx) - base) * scale);
                k = min(max(k, 0), ATE_NB - 1);
                atomicAdd(&S.bin[k + 1], 1);
            }
        }
        __syncthreads();
        {
            // inclusive scan of bin[1 .. NB]: 4 consecutive bins per thread, warp scan, warp totals in order
            constexpr int PER = ATE_NB / ATE_T;
            int v[PER], tot = 0;
#pragma unroll
            for (int k = 0; k < PER; ++k) { tot += S.bin[1 + tid * PER + k]; v[k] = tot; }
            int inc = tot;
#pragma unroll
            for (int off = 1; off < 32; off <<= 1) { const int y = __shfl_up_sync(GSF_FULL_MASK, inc, off); if (lane >= off) inc += y; }
            if (lane == 31) S.cursor[warp] = inc;
            __syncthreads();
            int pre = inc - tot;
            for (int w = 0; w < warp; ++w) pre += S.cursor[w];
            __syncthreads();
#pragma unroll
            for (int k = 0; k < PER; ++k) S.bin[1 + tid * PER + k] = pre + v[k];
        }
        __syncthreads();
        for (int k = tid; k < ATE_NB; k += ATE_T) S.cursor[k] = S.bin[k];
        __syncthreads();
        if (GSF_ATE_BBOX_STRIDE > 1) {
            m = S.bin[ATE_NB];                                  // the evaluation set's size: the histogram's total
            if (m == 0 || (!GLOBAL_WORK && m > A.cap)) {
                if (tid == 0) { o[0] = o[1] = o[2] = nan(""); o[3] = m > 0 ? -(double)m : 0.0; }
                continue;
            }
        }
#pragma unroll ATE_UNROLL
        for (int i = tid; i < n; i += ATE_T) {
            const double cx = gc[3 * (size_t)i], cy = gc[3 * (size_t)i + 1], cz = gc[3 * (size_t)i + 2];
            if (!row_has_nan(cx, cy, cz) && gts[i] > t0) {
                int k = (int)(((ax ? cy : cx) - base) * scale);
                k = min(max(k, 0), ATE_NB - 1);
                const int p = atomicAdd(&S.cursor[k], 1);
                cs[3 * (size_t)p] = cx; cs[3 * (size_t)p + 1] = cy; cs[3 * (size_t)p + 2] = cz;
            }
        }
        __syncthreads();

        // ---- queries in pose order: exact pruned nearest neighbour
        double v2[2] = {0.0, 0.0};
        unsigned long long kmn = ~0ull, kmx = 0ull;
#pragma unroll ATE_QUNROLL
        for (int i = tid; i < n; i += ATE_T) {
            const double cx = gc[3 * (size_t)i], cy = gc[3 * (size_t)i + 1], cz = gc[3 * (size_t)i + 2];
            if (row_has_nan(cx, cy, cz) || !(gts[i] > t0)) continue;
            const double px = gt[3 * (size_t)i], py = gt[3 * (size_t)i + 1], pz = gt[3 * (size_t)i + 2];
            double e;
            if (row_has_nan(px, py, pz)) { e = nan(""); S.any_nan = 1; }
            else {
                double dx = px - cx, dy = py - cy, dz = pz - cz;
                double best = dist2_rn(dx, dy, dz);               // own measurement first
                // every candidate nearer than the own measurement lies within r0 = sqrt(best) of the query along the binning axis:
                // the bins that this interval touches are one contiguous run of the bucketed array -- a single loop, no per-bin
                // edge tests.
                // (Keeping the candidates in registers across the four sweeps measured no faster: the re-reads hit L2.)
                const double qa = ax ? py : px;
                const double r0 = sqrt(best) + fmax(slack, 8.0 * 2.220446049250313e-16 * fabs(qa));      // (a query may lie outside the sampled box)
                int k0 = (int)fmax(0.0, fmin((qa - r0 - base) * scale, (double)(ATE_NB - 1)));
                int k1 = (int)fmax(0.0, fmin((qa + r0 - base) * scale, (double)(ATE_NB - 1)));
                if (!(r0 == r0)) { k0 = 0; k1 = ATE_NB - 1; }
                // no extra bin of margin: the bin index (v - base) * scale is a monotone function of v as computed, the same
                // expression places the candidates, and `slack` covers the rounding of sqrt and of qa -+ r0
                const int c1 = S.bin[k1 + 1];
                for (int c = S.bin[k0]; c < c1; ++c) {
                    dx = px - cs[3 * (size_t)c]; dy = py - cs[3 * (size_t)c + 1]; dz = pz - cs[3 * (size_t)c + 2];
                    best = fmin(best, dist2_rn(dx, dy, dz));
                }
                e = sqrt(best);
                const unsigned long long key = ate_key(e);
                kmn = min(kmn, key); kmx = max(kmx, key);
            }
            err[atomicAdd(&S.slot, 1)] = e;                            // selection does not depend on the order
            v2[0] += e; v2[1] += e * e;
        }
        block_sum<2>(v2, S.red);
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            kmn = min(kmn, __shfl_xor_sync(GSF_FULL_MASK, kmn, off)); kmx = max(kmx, __shfl_xor_sync(GSF_FULL_MASK, kmx, off));
        }
        if (lane == 0) { atomicMin(&S.kmin, kmn); atomicMax(&S.kmax, kmx); }
        __syncthreads();

        // ---- median: exact selection of rank (m - 1) / 2 and, for an even count, its upper neighbour (gsf_select.cuh)
        double med = nan("");
        if (!S.any_nan) {
            const double mu = v2[0] / m, var = v2[1] / m - mu * mu, sg = var > 0.0 ? sqrt(var) : 0.0;
            block_select<ATE_T, 4>(S.sel, err, m, (unsigned)((m - 1) / 2), !(m & 1), S.kmin, S.kmax, mu - sg, mu + sg);
            med = (m & 1) ? S.sel.v0 : 0.5 * (S.sel.v0 + S.sel.v1);
        }
        if (tid == 0) {
            o[0] = v2[0] / m; o[1] = med; o[2] = sqrt(v2[1] / m); o[3] = (double)m;
        }
    }
}

size_t ate_smem_bytes(int cap) { return ((sizeof(AteShared) + 15) & ~(size_t)15) + (size_t)cap * 32; }
int ate_smem_capacity(int max_smem) { return (int)(((size_t)max_smem - ((sizeof(AteShared) + 15) & ~(size_t)15)) / 32); }

cudaError_t launch_ate(const AteArgs& a, int num_sms, cudaStream_t stream) {
    if (a.B <= 0) return cudaSuccess;
    const bool global_work = a.work != nullptr;
    const size_t smem = global_work ? ate_smem_bytes(0) : ate_smem_bytes(a.cap);
    auto kern = global_work ? ate_nn_kernel<true> : ate_nn_kernel<false>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    if (e != cudaSuccess) return e;
    int per_sm = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, ATE_T, smem);
    if (e != cudaSuccess) return e;
    if (per_sm < 1) return cudaErrorInvalidConfiguration;
    long long grid = (long long)num_sms * per_sm;
    if (grid > a.B) grid = a.B;
    kern<<<(unsigned)grid, ATE_T, smem, stream>>>(a);
    return cudaGetLastError();
}

}  // namespace gsf
