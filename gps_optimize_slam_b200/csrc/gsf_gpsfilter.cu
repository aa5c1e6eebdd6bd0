// N3: GNSS outlier pre-filter, the per-window polynomial RANSAC of filter_gps_outliers_ransac
// (/root/reference/EKFGPSSLAM.py:136-247): for every window and coordinate axis the reference fits
//   make_pipeline(PolynomialFeatures(degree), RANSACRegressor(min_samples, residual_threshold, max_trials))
// (sklearn: LinearRegression base estimator, absolute-error loss) to (t, x_axis) and keeps the points that are inliers
// on all three axes in at least one window (:216-219).  The RANSAC loop restated here is sklearn's
// (linear_model/_ransac.py, RANSACRegressor.fit): per trial a least-squares polynomial through `min_samples` sampled
// points, residuals of all window points, inliers = residual <= threshold; a trial with fewer inliers than the best so far
// is skipped; on a tie the R^2 score of the sample model on its inliers decides (a worse score is skipped); after every
// accepted trial max_trials shrinks to _dynamic_max_trials(n_inliers_best, n, min_samples, 0.99).  The sample indices are
// HOST-SUPPLIED (what sklearn draws with sample_without_replacement from numpy's global RNG), and so is the table of
// _dynamic_max_trials values (indexed by the inlier count), so that a seeded reference run and this kernel execute the
// same trials and stop after the same number of them; `n_trials` tells the host how many draws the fit consumed.
// One warp per fit (window x axis); the fit itself is tiny (normal equations of degree <= 3 in centred, scaled time --
// the reference's uncentred Vandermonde columns are the same polynomial space), the residual pass runs over the lanes.
#include "gsf_common.cuh"
#include "gsf_internal.cuh"

namespace gsf {

constexpr int PR_MAXDEG = 3;

struct PolyModel { double c[PR_MAXDEG + 1]; double t0, inv_s; };

__device__ __forceinline__ double poly_eval(const PolyModel& M, int deg, double t) {
    const double u = (t - M.t0) * M.inv_s;
    double p = M.c[deg];
    for (int k = deg - 1; k >= 0; --k) p = fma(p, u, M.c[k]);
    return p;
}

// least squares of degree `deg` through the ms sampled points (every lane computes the same model)
__device__ bool poly_fit(const double* __restrict__ t, const double* __restrict__ y, int ystride, const int* __restrict__ widx,
                         const int* __restrict__ smp, int ms, int deg, PolyModel& M) {
    double tm = 0.0;
    for (int j = 0; j < ms; ++j) tm += t[widx[smp[j]]];
    tm /= (double)ms;
    double sc = 0.0;
    for (int j = 0; j < ms; ++j) sc = fmax(sc, fabs(t[widx[smp[j]]] - tm));
    if (!(sc > 0.0)) sc = 1.0;
    M.t0 = tm; M.inv_s = 1.0 / sc;
    double S[2 * PR_MAXDEG + 1], B[PR_MAXDEG + 1];
    for (int k = 0; k <= 2 * deg; ++k) S[k] = 0.0;
    for (int k = 0; k <= deg; ++k) B[k] = 0.0;
    const double y0 = y[(size_t)widx[smp[0]] * ystride];                 // pivot: sums of differences, not of 5e6 m coordinates
    for (int j = 0; j < ms; ++j) {
        const int g = widx[smp[j]];
        const double u = (t[g] - tm) * M.inv_s, v = y[(size_t)g * ystride] - y0;
        double p = 1.0;
        for (int k = 0; k <= 2 * deg; ++k) { S[k] += p; if (k <= deg) B[k] += v * p; p *= u; }
    }
    // normal equations A c = B, A[r][q] = S[r + q]: Gaussian elimination with partial pivoting (<= 4 x 4)
    double A[PR_MAXDEG + 1][PR_MAXDEG + 2];
    const int n = deg + 1;
    for (int r = 0; r < n; ++r) { for (int q = 0; q < n; ++q) A[r][q] = S[r + q]; A[r][n] = B[r]; }
    bool ok = true;
    for (int col = 0; col < n; ++col) {
        int piv = col; double big = fabs(A[col][col]);
        for (int r = col + 1; r < n; ++r) if (fabs(A[r][col]) > big) { big = fabs(A[r][col]); piv = r; }
        if (!(big > 1e-14 * (double)ms)) { ok = false; break; }          // rank-deficient sample (repeated times): the reference's
        if (piv != col) for (int q = col; q <= n; ++q) { const double tmp = A[col][q]; A[col][q] = A[piv][q]; A[piv][q] = tmp; }   // min-norm fit is not reproduced
        const double inv = 1.0 / A[col][col];
        for (int r = col + 1; r < n; ++r) {
            const double f = A[r][col] * inv;
            for (int q = col; q <= n; ++q) A[r][q] -= f * A[col][q];
        }
    }
    if (ok) {
        for (int r = n - 1; r >= 0; --r) {
            double acc = A[r][n];
            for (int q = r + 1; q < n; ++q) acc -= A[r][q] * M.c[q];
            M.c[r] = acc / A[r][r];
        }
        M.c[0] += y0;
    } else {
        for (int r = 0; r < n; ++r) M.c[r] = 0.0;
        double mean = 0.0;
        for (int j = 0; j < ms; ++j) mean += y[(size_t)widx[smp[j]] * ystride] - y0;
        M.c[0] = y0 + mean / (double)ms;                                 // constant model (what a rank-1 design leaves)
    }
    return ok;
}

struct PolyRansacArgs {
    const double* t; const double* y; int ystride;
    const int* widx; const long long* fit_off; const int* fit_axis;      // fit f: window indices widx[fit_off[f] .. fit_off[f+1]), column fit_axis[f] of y
    const int* samples;                                                  // [F, max_trials, ms] indices into the fit's window
    const int* dyn_trials; const long long* dyn_off;                     // dyn_trials[dyn_off[f] + n_inliers] = _dynamic_max_trials(...)
    int F, ms, deg, max_trials; double thr;
    unsigned char* mask;                                                 // [fit_off[F]] inlier mask of the best trial
    int* n_trials; int* status;                                          // status 1: no consensus set (sklearn raises ValueError)
};

__global__ void __launch_bounds__(128) poly_ransac_kernel(const PolyRansacArgs A) {
    const int lane = threadIdx.x & 31;
    const int f = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (f >= A.F) return;
    const long long o0 = A.fit_off[f];
    const int nw = (int)(A.fit_off[f + 1] - o0);
    const int* __restrict__ widx = A.widx + o0;
    const double* __restrict__ y = A.y + A.fit_axis[f];
    const int* __restrict__ dyn = A.dyn_trials + A.dyn_off[f];
    int n_best = 1, trials = 0, max_trials = A.max_trials;
    double score_best = -INFINITY;
    bool have = false;
    PolyModel best{};
    while (trials < max_trials) {
        const int* smp = A.samples + ((size_t)f * A.max_trials + trials) * A.ms;
        ++trials;
        PolyModel M;
        poly_fit(A.t, y, A.ystride, widx, smp, A.ms, A.deg, M);
        // residuals of all window points, inlier count
        int cnt = 0;
        for (int k = lane; k < nw; k += 32) {
            const int g = widx[k];
            if (fabs(y[(size_t)g * A.ystride] - poly_eval(M, A.deg, A.t[g])) <= A.thr) ++cnt;
        }
        for (int ofs = 16; ofs > 0; ofs >>= 1) cnt += __shfl_xor_sync(GSF_FULL_MASK, cnt, ofs);
        if (cnt < n_best) continue;                                      // fewer inliers: skip
        // R^2 of the sample model on its inliers (pivot-shifted sums)
        const double y0 = y[(size_t)widx[0] * A.ystride];
        double sr = 0.0, s1 = 0.0, s2 = 0.0;
        for (int k = lane; k < nw; k += 32) {
            const int g = widx[k];
            const double yy = y[(size_t)g * A.ystride], r = yy - poly_eval(M, A.deg, A.t[g]);
            if (fabs(r) <= A.thr) { sr += r * r; const double v = yy - y0; s1 += v; s2 += v * v; }
        }
        for (int ofs = 16; ofs > 0; ofs >>= 1) {
            sr += __shfl_xor_sync(GSF_FULL_MASK, sr, ofs); s1 += __shfl_xor_sync(GSF_FULL_MASK, s1, ofs); s2 += __shfl_xor_sync(GSF_FULL_MASK, s2, ofs);
        }
        const double sst = s2 - s1 * s1 / (double)cnt;
        const double score = sst > 0.0 ? 1.0 - sr / sst : (sr == 0.0 ? 1.0 : 0.0);      // r2_score's zero-variance convention
        if (cnt == n_best && score < score_best) continue;               // tie with a worse score: skip
        n_best = cnt; score_best = score; best = M; have = true;
        max_trials = min(max_trials, dyn[cnt]);
    }
    if (lane == 0) { A.n_trials[f] = trials; A.status[f] = have ? 0 : 1; }
    for (int k = lane; k < nw; k += 32) {
        const int g = widx[k];
        A.mask[o0 + k] = (have && fabs(y[(size_t)g * A.ystride] - poly_eval(best, A.deg, A.t[g])) <= A.thr) ? 1 : 0;
    }
}

cudaError_t launch_poly_ransac(const double* t, const double* y, int ystride, const int* widx, const long long* fit_off, const int* fit_axis,
                               const int* samples, const int* dyn_trials, const long long* dyn_off, int F, int ms, int deg, int max_trials,
                               double thr, unsigned char* mask, int* n_trials, int* status, cudaStream_t stream) {
    if (F <= 0) return cudaSuccess;
    if (deg < 0 || deg > PR_MAXDEG || ms < deg + 1) return cudaErrorInvalidValue;
    PolyRansacArgs a;
    a.t = t; a.y = y; a.ystride = ystride; a.widx = widx; a.fit_off = fit_off; a.fit_axis = fit_axis; a.samples = samples;
    a.dyn_trials = dyn_trials; a.dyn_off = dyn_off; a.F = F; a.ms = ms; a.deg = deg; a.max_trials = max_trials; a.thr = thr;
    a.mask = mask; a.n_trials = n_trials; a.status = status;
    poly_ransac_kernel<<<(F + 3) / 4, 128, 0, stream>>>(a);
    return cudaGetLastError();
}

}  // namespace gsf
