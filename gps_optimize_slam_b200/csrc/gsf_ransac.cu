// gsf_sim3_ransac: compute_sim3_transform_robust (EKFGPSSLAM.py:389-426) with HOST-SUPPLIED sample
// indices, so that a seeded reference run (np.random.seed(k) before the call) and this path evaluate
// exactly the same trials:
//   for each trial: Umeyama on the `m` sampled pairs (:409-410), residual norms of ALL points under
//   that transform (:412-413), inlier count (:414-415); the first trial with the strictly largest
//   count wins (:416-417); its inlier mask is the point set of the final fit (:422-423).
// One thread block per trial (trials are independent: grid-stride over them), the sampled fit done by
// one thread (m is 4 in the shipped CONFIG), the residual sweep by the whole block; a single-block
// selection kernel picks the winner and rebuilds its mask; the final fit reuses the tiled Umeyama
// kernels (gsf_kernels.cu) with that mask.  Deterministic: integer counts, fixed tie rule.
#include "gsf_common.cuh"
#include "gsf_internal.cuh"

namespace gsf {

constexpr int RANSAC_THREADS = 128;
constexpr int TRIAL_STRIDE = 14;     // R(9) t(3) s status

__device__ __forceinline__ bool sim3_inlier(const double* __restrict__ T, const double* __restrict__ src, const double* __restrict__ dst,
                                            long long i, double thr) {
    const double x = src[3 * i], y = src[3 * i + 1], z = src[3 * i + 2];
    const double s = T[12];
    const double d0 = s * (T[0] * x + T[1] * y + T[2] * z) + T[9] - dst[3 * i];
    const double d1 = s * (T[3] * x + T[4] * y + T[5] * z) + T[10] - dst[3 * i + 1];
    const double d2 = s * (T[6] * x + T[7] * y + T[8] * z) + T[11] - dst[3 * i + 2];
    return sqrt(d0 * d0 + d1 * d1 + d2 * d2) < thr;         // np.linalg.norm(...) < residual_threshold; NaN -> false
}

__global__ void __launch_bounds__(RANSAC_THREADS) ransac_trials_kernel(const double* __restrict__ src, const double* __restrict__ dst, long long n,
                                                                       const int* __restrict__ samples, int m, int T, double thr,
                                                                       double* __restrict__ trials, int* __restrict__ counts) {
    __shared__ double tr[TRIAL_STRIDE];
    __shared__ int wcount[RANSAC_THREADS / 32];
    for (int trial = blockIdx.x; trial < T; trial += gridDim.x) {
        if (threadIdx.x == 0) {
            // compute_sim3_transform on the sample (:428-459): means, centred H, ss, then the shared SVD finish
            const int* idx = samples + (size_t)trial * m;
            double ms[3] = {0, 0, 0}, md[3] = {0, 0, 0};
            bool ok = m >= 3;
            for (int k = 0; k < m; ++k) {
                const long long i = idx[k];
                if (i < 0 || i >= n) { ok = false; break; }
                for (int c = 0; c < 3; ++c) { ms[c] += src[3 * i + c]; md[c] += dst[3 * i + c]; }
            }
            int st = ST_TOO_FEW_POINTS;
            double R[9], t[3], s = 1.0;
            if (ok) {
                const double inv = 1.0 / (double)m;
                for (int c = 0; c < 3; ++c) { ms[c] *= inv; md[c] *= inv; }
                double H[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0}, ss = 0.0;
                for (int k = 0; k < m; ++k) {
                    const long long i = idx[k];
                    const double a[3] = {src[3 * i] - ms[0], src[3 * i + 1] - ms[1], src[3 * i + 2] - ms[2]};
                    const double b[3] = {dst[3 * i] - md[0], dst[3 * i + 1] - md[1], dst[3 * i + 2] - md[2]};
                    for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) H[3 * r + c] += a[r] * b[c];
                    ss += a[0] * a[0] + a[1] * a[1] + a[2] * a[2];
                }
                st = umeyama_finish(m, ms, md, H, ss, R, t, s);
                if (!(ss == ss)) st |= ST_TOO_FEW_POINTS;          // NaN sample: numpy's SVD raises -> trial skipped (:411)
            }
            for (int k = 0; k < 9; ++k) tr[k] = R[k];
            tr[9] = t[0]; tr[10] = t[1]; tr[11] = t[2]; tr[12] = s; tr[13] = (double)st;
        }
        __syncthreads();
        const bool valid = !(((int)tr[13]) & ST_TOO_FEW_POINTS);
        int cnt = 0;
        if (valid)
            for (long long i = threadIdx.x; i < n; i += RANSAC_THREADS) cnt += sim3_inlier(tr, src, dst, i, thr) ? 1 : 0;
        cnt = warp_sum_i(cnt);
        if ((threadIdx.x & 31) == 0) wcount[threadIdx.x >> 5] = cnt;
        __syncthreads();
        if (threadIdx.x == 0) {
            int total = 0;
            for (int w = 0; w < RANSAC_THREADS / 32; ++w) total += wcount[w];
            counts[trial] = valid ? total : -1;
            for (int k = 0; k < TRIAL_STRIDE; ++k) trials[(size_t)trial * TRIAL_STRIDE + k] = tr[k];
        }
        __syncthreads();
    }
}

// Winner = first trial with the strictly largest count (:416); rebuild its inlier mask; info = {max_inliers, trial}.
__global__ void __launch_bounds__(256) ransac_select_kernel(const double* __restrict__ src, const double* __restrict__ dst, long long n,
                                                            const double* __restrict__ trials, const int* __restrict__ counts, int T,
                                                            double thr, int min_inliers, unsigned char* __restrict__ mask,
                                                            long long* __restrict__ offsets2, int* __restrict__ info) {
    __shared__ int bc[256], bi[256];
    __shared__ double tr[TRIAL_STRIDE];
    int best = -1, bidx = 0x7fffffff;
    for (int k = threadIdx.x; k < T; k += 256) {
        const int c = counts[k];
        if (c > best) { best = c; bidx = k; }               // k increases: the first maximum is kept
    }
    bc[threadIdx.x] = best; bi[threadIdx.x] = bidx;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) {
            const int c2 = bc[threadIdx.x + o], i2 = bi[threadIdx.x + o];
            if (c2 > bc[threadIdx.x] || (c2 == bc[threadIdx.x] && i2 < bi[threadIdx.x])) { bc[threadIdx.x] = c2; bi[threadIdx.x] = i2; }
        }
        __syncthreads();
    }
    best = bc[0]; bidx = bi[0];
    const bool usable = best >= 0 && best >= min_inliers;   // :419
    if (threadIdx.x < TRIAL_STRIDE && best >= 0) tr[threadIdx.x] = trials[(size_t)bidx * TRIAL_STRIDE + threadIdx.x];
    if (threadIdx.x == 0) { info[0] = best; info[1] = best >= 0 ? bidx : -1; info[2] = usable ? 1 : 0; offsets2[0] = 0; offsets2[1] = n; }
    __syncthreads();
    for (long long i = threadIdx.x; i < n; i += 256) mask[i] = (usable && sim3_inlier(tr, src, dst, i, thr)) ? 1 : 0;
}

cudaError_t launch_sim3_ransac(const double* src, const double* dst, long long n, const int* samples, int m, int T, double thr,
                               int min_inliers, double* work, unsigned char* mask, double* R, double* t, double* s,
                               int* info, int* status, int num_sms, cudaStream_t stream) {
    // work layout: trials [T*14] | counts [T ints, padded to doubles] | offsets2 [2 long long] | Umeyama tile stats
    double* trials = work;
    int* counts = reinterpret_cast<int*>(trials + (size_t)T * TRIAL_STRIDE);
    long long* offsets2 = reinterpret_cast<long long*>(trials + (size_t)T * TRIAL_STRIDE + (T + 1) / 2);
    double* uwork = reinterpret_cast<double*>(offsets2 + 2);
    int grid = T < num_sms * 8 ? T : num_sms * 8;
    if (grid < 1) grid = 1;
    ransac_trials_kernel<<<grid, RANSAC_THREADS, 0, stream>>>(src, dst, n, samples, m, T, thr, trials, counts);
    ransac_select_kernel<<<1, 256, 0, stream>>>(src, dst, n, trials, counts, T, thr, min_inliers, mask, offsets2, info);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    // final fit on the winner's inliers (:422-423); an all-zero mask yields ST_TOO_FEW_POINTS = (None, None, None)
    return launch_umeyama(src, dst, offsets2, mask, 1, n, uwork, R, t, s, status, stream);
}
long long sim3_ransac_work_doubles(int T, long long n) {
    return (long long)T * TRIAL_STRIDE + (T + 1) / 2 + 2 + (long long)sim3_tiles_for(n) * 20;
}

}  // namespace gsf
