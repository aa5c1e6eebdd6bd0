// Exact order statistics of m non-negative doubles held in shared memory, by a whole thread block: the median step of
// the evaluation (np.median, /root/reference/EKFGPSSLAM.py:1033) for the NN-ATE kernel and the noise-grid combine kernel.
//
// Linear bins between two bracket values, one histogram pass, a block scan over the bins, then the members of the bin that
// holds the wanted rank are gathered and ranked by one warp with shuffles.  The bin index is a monotone function of the
// value, so order statistics are exact whatever the bracket; a good bracket only makes one level enough: callers pass
// [mean - sigma, mean + sigma] (the median of any sample lies in it), which puts a few elements in each of the ~1000
// bins, where the 8-bit radix digits of the IEEE patterns put thousands in a handful (shared-memory atomics on the
// same address serialise: the radix version of this step cost 3x the nearest-neighbour search).  If the bin is still
// crowded (> 32 members: ties, or a degenerate bracket) the interval is narrowed to the bin's own key range and the
// level repeats; identical keys end it at once.
#pragma once
#include "gsf_common.cuh"

namespace gsf {

template <int THREADS, int BPT>
struct SelectShared {
    unsigned int hist[THREADS * BPT];
    unsigned int wtot[32];
    unsigned long long wk[2][32];
    unsigned long long list[32];
    unsigned int rank, cnt, bin, lcount, le;
    unsigned long long above;
    double v0, v1;
    int have_v1;
};

__device__ __forceinline__ int select_bin(double e, double a, double scale, int nbins) {
    const double t = (e - a) * scale;
    return t < (double)(nbins - 1) ? (t > 0.0 ? (int)t : 0) : nbins - 1;
}

// Value of rank r (0-based) -> S.v0 and, if want_next, of rank r + 1 -> S.v1.  [klo, khi]: smallest / largest key present;
// [blo, bhi]: bracket of the first level.  Every thread of the block must call; the results are valid after return.
template <int THREADS, int BPT>
__device__ void block_select(SelectShared<THREADS, BPT>& S, const double* __restrict__ err, int m, unsigned r, bool want_next,
                             unsigned long long klo, unsigned long long khi, double blo, double bhi) {
    constexpr int NBINS = THREADS * BPT;
    constexpr int NWARP = THREADS / 32;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    unsigned long long ka = klo, kb = khi;
    unsigned rank = r;
    unsigned below = 0;                                    // elements under the current interval (for the rank r + 1 shortcut)
    if (tid == 0) S.have_v1 = 0;
    for (int level = 0; level < 16; ++level) {
        if (ka == kb) {                                    // every remaining element is the same value
            if (tid == 0) { S.v0 = __longlong_as_double((long long)ka); }
            __syncthreads();
            break;
        }
        double a = __longlong_as_double((long long)ka), b = __longlong_as_double((long long)kb);
        if (level == 0 && blo < bhi) { a = fmax(a, blo); b = fmin(b, bhi); if (!(a < b)) { a = __longlong_as_double((long long)ka); b = __longlong_as_double((long long)kb); } }
        const double scale = (double)NBINS / (b - a);
#pragma unroll
        for (int q = 0; q < BPT; ++q) S.hist[tid + q * THREADS] = 0u;
        if (tid == 0) S.lcount = 0u;
        __syncthreads();
        for (int k = tid; k < m; k += THREADS) {
            const double e = err[k];
            const unsigned long long key = (unsigned long long)__double_as_longlong(e);
            if (key >= ka && key <= kb) atomicAdd(&S.hist[select_bin(e, a, scale, NBINS)], 1u);
        }
        __syncthreads();
        // exclusive scan over the bins: BPT consecutive bins per thread
        unsigned c[BPT], tot = 0;
#pragma unroll
        for (int q = 0; q < BPT; ++q) { c[q] = S.hist[tid * BPT + q]; tot += c[q]; }
        unsigned inc = tot;
#pragma unroll
        for (int ofs = 1; ofs < 32; ofs <<= 1) { const unsigned y = __shfl_up_sync(GSF_FULL_MASK, inc, ofs); if (lane >= ofs) inc += y; }
        if (lane == 31) S.wtot[warp] = inc;
        __syncthreads();
        unsigned pre = 0;
        {
            const unsigned w = lane < NWARP ? S.wtot[lane] : 0u;       // warp totals: prefix by one more shuffle scan
            unsigned wi = w;
#pragma unroll
            for (int ofs = 1; ofs < 32; ofs <<= 1) { const unsigned y = __shfl_up_sync(GSF_FULL_MASK, wi, ofs); if (lane >= ofs) wi += y; }
            pre = __shfl_sync(GSF_FULL_MASK, wi - w, warp);
        }
        unsigned exc = pre + inc - tot;
        if (tot > 0 && rank >= exc && rank < exc + tot) {
#pragma unroll
            for (int q = 0; q < BPT; ++q) {
                if (rank >= exc && rank < exc + c[q]) { S.rank = rank - exc; S.cnt = c[q]; S.bin = (unsigned)(tid * BPT + q); S.le = below + exc + c[q]; }
                exc += c[q];
            }
        }
        __syncthreads();
        const int bin = (int)S.bin;
        const unsigned cnt = S.cnt;
        rank = S.rank;
        if (cnt <= 32u) {
            // gather the members of the bin; warp 0 ranks them with shuffles
            for (int k = tid; k < m; k += THREADS) {
                const double e = err[k];
                const unsigned long long key = (unsigned long long)__double_as_longlong(e);
                if (key >= ka && key <= kb && select_bin(e, a, scale, NBINS) == bin) S.list[atomicAdd(&S.lcount, 1u)] = key;
            }
            __syncthreads();
            if (warp == 0) {
                const unsigned long long xk = lane < (int)cnt ? S.list[lane] : ~0ull;
                unsigned less = 0;
                for (unsigned q = 0; q < cnt; ++q) {
                    const unsigned long long yk = __shfl_sync(GSF_FULL_MASK, xk, (int)q);
                    less += (yk < xk || (yk == xk && q < (unsigned)lane)) ? 1u : 0u;
                }
                if (lane < (int)cnt && less == rank) S.v0 = __longlong_as_double((long long)xk);
                if (lane < (int)cnt && less == rank + 1u) { S.v1 = __longlong_as_double((long long)xk); S.have_v1 = 1; }
            }
            __syncthreads();
            break;
        }
        // crowded bin: narrow the interval to its members' key range
        below = S.le - cnt;
        unsigned long long lo2 = ~0ull, hi2 = 0ull;
        for (int k = tid; k < m; k += THREADS) {
            const double e = err[k];
            const unsigned long long key = (unsigned long long)__double_as_longlong(e);
            if (key >= ka && key <= kb && select_bin(e, a, scale, NBINS) == bin) { lo2 = min(lo2, key); hi2 = max(hi2, key); }
        }
#pragma unroll
        for (int ofs = 16; ofs > 0; ofs >>= 1) { lo2 = min(lo2, __shfl_xor_sync(GSF_FULL_MASK, lo2, ofs)); hi2 = max(hi2, __shfl_xor_sync(GSF_FULL_MASK, hi2, ofs)); }
        if (lane == 0) { S.wk[0][warp] = lo2; S.wk[1][warp] = hi2; }
        __syncthreads();
        lo2 = lane < NWARP ? S.wk[0][lane] : ~0ull; hi2 = lane < NWARP ? S.wk[1][lane] : 0ull;
#pragma unroll
        for (int ofs = 16; ofs > 0; ofs >>= 1) { lo2 = min(lo2, __shfl_xor_sync(GSF_FULL_MASK, lo2, ofs)); hi2 = max(hi2, __shfl_xor_sync(GSF_FULL_MASK, hi2, ofs)); }
        ka = lo2; kb = hi2;
        __syncthreads();
    }
    if (want_next && !S.have_v1) {
        // rank r + 1: the same value if enough elements are <= it, else the smallest element above it
        const double v0 = S.v0;
        const unsigned long long k0 = (unsigned long long)__double_as_longlong(v0);
        unsigned le = 0; unsigned long long ab = ~0ull;
        for (int k = tid; k < m; k += THREADS) {
            const unsigned long long key = (unsigned long long)__double_as_longlong(err[k]);
            if (key <= k0) ++le; else ab = min(ab, key);
        }
#pragma unroll
        for (int ofs = 16; ofs > 0; ofs >>= 1) { le += __shfl_xor_sync(GSF_FULL_MASK, le, ofs); ab = min(ab, __shfl_xor_sync(GSF_FULL_MASK, ab, ofs)); }
        if (lane == 0) { S.wtot[warp] = le; S.wk[0][warp] = ab; }
        __syncthreads();
        if (warp == 0) {
            le = lane < NWARP ? S.wtot[lane] : 0u; ab = lane < NWARP ? S.wk[0][lane] : ~0ull;
#pragma unroll
            for (int ofs = 16; ofs > 0; ofs >>= 1) { le += __shfl_xor_sync(GSF_FULL_MASK, le, ofs); ab = min(ab, __shfl_xor_sync(GSF_FULL_MASK, ab, ofs)); }
            if (lane == 0) S.v1 = le > r + 1u ? v0 : __longlong_as_double((long long)ab);
        }
        __syncthreads();
    }
}

}  // namespace gsf
