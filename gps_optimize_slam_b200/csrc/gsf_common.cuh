// Device-side building blocks shared by every kernel of the GPS/SLAM fusion path.
// fp64 throughout; quaternions are xyzw (scalar last) with the Hamilton product,
// exactly as scipy's Rotation (the reference's rotation library) defines them.
#pragma once
#include <stdint.h>
#include <math.h>
#ifdef __CUDACC__
#include <cuda_runtime.h>
#define GSF_HD __host__ __device__
#define GSF_D __device__
#else
// Host-only build (tests/hostmath): the scalar math below is compiled with g++ so that the
// CPU test-suite can check it against numpy/scipy without a GPU.
#define GSF_HD
#define GSF_D
#endif

#define GSF_FULL_MASK 0xffffffffu
#ifndef __CUDACC__
#define __forceinline__ inline
#endif

namespace gsf {

// ----------------------------------------------------------------------------- status codes
// Per-trajectory status written by the batched kernels (mirrors the reference's
// None-triples / ValueError / RuntimeError exits, EKFGPSSLAM.py:431, :975, :997, :1003).
enum : int {
    ST_OK = 0,
    ST_TOO_FEW_POINTS = 1,     // < 3 points for Umeyama (:431) / < min_samples for Sim3 (:975,:997)
    ST_DEGENERATE = 2,         // rank-deficient cross-covariance: rotation not unique
    ST_BAD_QUATERNION = 4,     // zero-norm SLAM quaternion met (scipy raises ValueError)
    ST_EMPTY = 8,              // zero-length trajectory
    ST_RANSAC_OUTLIERS = 16,   // all-points fit leaves residuals >= threshold: the reference's
                               // unseeded RANSAC could pick a different inlier set here
    ST_TOO_LONG = 32,          // trajectory does not fit the shared-memory staging buffer
    ST_GRID_NEEDS_ALL_VALID = 64,  // hypothesis grid: the trajectory has poses without GNSS (outage/RTS unsupported there)
    ST_NEEDS_FP64 = 128,       // fp32 mode: the trajectory needs the general machinery (outage / gap / window / ...): use the fp64 path
    ST_DEFERRED = 1 << 30,     // internal: left by the fast kernel for the general kernel (never returned)
};

// EKF / pipeline parameters for one trajectory (or shared by the batch).  Values are the
// reference's CONFIG entries (EKFGPSSLAM.py:22-71) flattened.
struct FuseParams {
    double p0[7];            // ekf.initial_cov_diag
    double q[7];             // ekf.process_noise_diag (per second)
    double r[3];             // ekf.meas_noise_diag (used as variances, :686)
    double gap_threshold;    // time_alignment.max_gps_gap_threshold
    double max_duration;     // sim3_ransac.max_initial_duration
    double yaw_rate_thresh;  // rts_decision threshold in rad/s
    double residual_thresh;  // sim3_ransac.residual_threshold
    double eval_skip;        // 5.0 s evaluation cut (:1021)
    int min_samples;         // sim3_ransac.min_samples
    int sharp_turn_steps;    // rts_decision.default_ekf_transition_steps_on_sharp_turn
};
static_assert(sizeof(FuseParams) == 23 * 8, "FuseParams layout is part of the C ABI");

// ----------------------------------------------------------------------------- small math
struct Quat { double x, y, z, w; };

#ifdef __CUDACC__
// library rsqrt kept out of line: its special-case code would otherwise be inlined at every call site of the hot
// Jacobi chain (instruction-cache footprint)
static __device__ __noinline__ double rsqrt_lib(double x) { return rsqrt(x); }
#endif
GSF_HD __forceinline__ double rsqrt_(double x) {
#ifdef __CUDA_ARCH__
    // CUDA's rsqrt() arithmetic (hardware seed + one third-order step) for arguments away from the ends of the
    // exponent range; the library call keeps the special cases (0, denormal, inf, NaN, |exponent| > 1000)
    const unsigned ex = ((unsigned)__double2hiint(x) >> 20) - 24u;      // sign bit set -> huge -> library path
    if (ex < 2000u) {
        double r;
        asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
        const double e = fma(-x, r * r, 1.0);
        return fma(fma(e, 0.375, 0.5), r * e, r);
    }
    return rsqrt_lib(x);
#else
    return 1.0 / sqrt(x);
#endif
}

// 1/x for finite positive normal x.  Device: hardware seed (rel. error 2^-23) + two Newton
// steps (<= 1 ulp, 5 instructions instead of the ~10-instruction IEEE division); host: 1.0/x.
GSF_HD __forceinline__ double rcp_(double x) {
#ifdef __CUDA_ARCH__
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    double e = fma(-x, r, 1.0);
    r = fma(r, e, r);
    e = fma(-x, r, 1.0);
    return fma(r, e, r);
#else
    return 1.0 / x;
#endif
}

GSF_HD __forceinline__ Quat qmul(const Quat& p, const Quat& q) {
    // scipy compose_quat(p, q): rotation q applied first, then p.
    Quat r;
    // one multiply + three FMAs per component (16 operations; the sum-of-products form compiled to 19)
    r.x = fma(p.w, q.x, fma(q.w, p.x, fma(p.y, q.z, -(p.z * q.y))));
    r.y = fma(p.w, q.y, fma(q.w, p.y, fma(p.z, q.x, -(p.x * q.z))));
    r.z = fma(p.w, q.z, fma(q.w, p.z, fma(p.x, q.y, -(p.y * q.x))));
    r.w = fma(p.w, q.w, -fma(p.x, q.x, fma(p.y, q.y, p.z * q.z)));
    return r;
}
GSF_HD __forceinline__ Quat qconj(const Quat& q) { return Quat{-q.x, -q.y, -q.z, q.w}; }
GSF_HD __forceinline__ double qnorm2(const Quat& q) { return q.x * q.x + q.y * q.y + q.z * q.z + q.w * q.w; }
GSF_HD __forceinline__ Quat qscale(const Quat& q, double s) { return Quat{q.x * s, q.y * s, q.z * s, q.w * s}; }

// scipy Rotation.from_quat normalisation (raises on zero norm: caller checks n2 == 0 first).
GSF_HD __forceinline__ Quat qunit(const Quat& q) { return qscale(q, rsqrt_(qnorm2(q))); }

// ExtendedKalmanFilter.normalize_quaternion (EKFGPSSLAM.py:697-700): identity below 1e-9.
GSF_HD __forceinline__ Quat qunit_or_identity(const Quat& q) {
    double n2 = qnorm2(q);
    if (!(n2 > 1e-18)) return Quat{0.0, 0.0, 0.0, 1.0};
    return qscale(q, rsqrt_(n2));
}

// scipy as_matrix (row-major).  Homogeneous of degree 2 in q; q is unit on every call site.
GSF_HD __forceinline__ void qmat(const Quat& q, double* M) {
    double x2 = q.x * q.x, y2 = q.y * q.y, z2 = q.z * q.z, w2 = q.w * q.w;
    double xy = q.x * q.y, zw = q.z * q.w, xz = q.x * q.z, yw = q.y * q.w, yz = q.y * q.z, xw = q.x * q.w;
    M[0] = x2 - y2 - z2 + w2; M[1] = 2.0 * (xy - zw);     M[2] = 2.0 * (xz + yw);
    M[3] = 2.0 * (xy + zw);     M[4] = -x2 + y2 - z2 + w2; M[5] = 2.0 * (yz - xw);
    M[6] = 2.0 * (xz - yw);     M[7] = 2.0 * (yz + xw);    M[8] = -x2 - y2 + z2 + w2;
}
GSF_HD __forceinline__ void mat_vec(const double* M, double x, double y, double z, double& ox, double& oy, double& oz) {
    ox = M[0] * x + M[1] * y + M[2] * z;
    oy = M[3] * x + M[4] * y + M[5] * z;
    oz = M[6] * x + M[7] * y + M[8] * z;
}
GSF_HD __forceinline__ void matT_vec(const double* M, double x, double y, double z, double& ox, double& oy, double& oz) {
    ox = M[0] * x + M[3] * y + M[6] * z;
    oy = M[1] * x + M[4] * y + M[7] * z;
    oz = M[2] * x + M[5] * y + M[8] * z;
}

// scipy _from_matrix_orthogonal: pick the largest of (m00, m11, m22, trace), no sign
// canonicalisation, normalise.
GSF_HD __forceinline__ Quat quat_from_matrix(const double* m) {
    double tr = m[0] + m[4] + m[8];
    int choice = 0; double best = m[0];
    if (m[4] > best) { best = m[4]; choice = 1; }
    if (m[8] > best) { best = m[8]; choice = 2; }
    if (tr > best) { choice = 3; }
    Quat q;
    if (choice == 0)      { q.x = 1.0 - tr + 2.0 * m[0]; q.y = m[3] + m[1]; q.z = m[6] + m[2]; q.w = m[7] - m[5]; }
    else if (choice == 1) { q.x = m[3] + m[1]; q.y = 1.0 - tr + 2.0 * m[4]; q.z = m[7] + m[5]; q.w = m[2] - m[6]; }
    else if (choice == 2) { q.x = m[6] + m[2]; q.y = m[7] + m[5]; q.z = 1.0 - tr + 2.0 * m[8]; q.w = m[3] - m[1]; }
    else                  { q.x = m[7] - m[5]; q.y = m[2] - m[6]; q.z = m[3] - m[1]; q.w = 1.0 + tr; }
    return qunit(q);
}

// First angle of scipy as_euler('zyx') (extrinsic z-y-x; i=2, j=1, k=0, sign=-1), including
// its gimbal-lock cases (_get_angles, eps = 1e-7).  Used by the sharp-turn gate
// (EKFGPSSLAM.py:819-820).
GSF_HD inline double yaw_zyx(const Quat& q) {
    const double PI = 3.141592653589793238462643383279502884;
    double a = q.w - q.y, b = q.z - q.x, c = q.y + q.w, d = -q.x - q.z;
    double half_sum = atan2(b, a), half_diff = atan2(d, c);
    double second = 2.0 * atan2(hypot(c, d), hypot(a, b));
    double first;
    if (fabs(second) <= 1e-7) first = 2.0 * half_sum;
    else if (fabs(second - PI) <= 1e-7) first = -2.0 * half_diff;
    else first = half_sum - half_diff;
    // (angle + pi) mod 2pi - pi with Python's sign-of-divisor modulo
    double m = fmod(first + PI, 2.0 * PI);
    if (m < 0.0) m += 2.0 * PI;
    return m - PI;
}

// One-sided (Hestenes) Jacobi SVD of a 3x3 matrix, columns kept in registers: high
// relative accuracy for the small singular directions, which decide the roll of a
// near-collinear track (KITTI-04: sigma = 3.4e6 / 10.7 / 1.45).
// One rotation of the column pair held in (a0, a1) / (v0, v1); returns true if it rotated.
GSF_HD __forceinline__ bool jacobi_rotate_pair(double* a0, double* a1, double* v0, double* v1) {
    const double alpha = a0[0] * a0[0] + a0[1] * a0[1] + a0[2] * a0[2];
    const double beta = a1[0] * a1[0] + a1[1] * a1[1] + a1[2] * a1[2];
    const double gamma = a0[0] * a1[0] + a0[1] * a1[1] + a0[2] * a1[2];
    if (gamma * gamma <= 1e-31 * alpha * beta || gamma == 0.0) return false;   // |cos angle| <= 3.2e-16 (1e-13 saves one
                                                                               // rotation in six and measured 0.7 % on the fused kernel: not taken)
    // rotation by theta in [-pi/4, pi/4] with tan(2 theta) = g2 / d:  cos(2 theta) = |d| / hyp, sin(2 theta) = sign(d) g2 / hyp,
    // c = sqrt((1 + cos 2theta) / 2), s = sin(2 theta) / (2 c)  -- two reciprocal square roots, no division
    // (dependent chain 23 operations instead of 30 for the tangent form; c^2 + s^2 = 1 to rounding either way)
    const double d = beta - alpha, g2 = 2.0 * gamma;
    const double w = d * d + g2 * g2;                 // > 0: gamma != 0 here
    const double rw = rsqrt_(w);
    const double c2 = fabs(d) * rw, s2 = (d >= 0.0 ? g2 : -g2) * rw;
    const double cc = 0.5 + 0.5 * c2;                 // in [0.5, 1]
    const double rc = rsqrt_(cc);
    const double c = cc * rc, sn = 0.5 * s2 * rc;
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        const double x = a0[r], y = a1[r];
        a0[r] = c * x - sn * y; a1[r] = sn * x + c * y;
        const double vx = v0[r], vy = v1[r];
        v0[r] = c * vx - sn * vy; v1[r] = sn * vx + c * vy;
    }
    return true;
}

// Umeyama's rotation + scale from the (un-normalised) cross-covariance H = src_c^T dst_c,
// restating EKFGPSSLAM.py:439-449:  U,S,Vt = svd(H); R = Vt^T U^T; if det(R) < 0 negate the
// last row of Vt; scale numerator = S1 + S2 + S3*det(R_fixed) = S1+S2+S3 in both branches.
// With singular pairs (u_i, v_i) sorted by sigma, the det=+1 matrix V diag(1,1,d) U^T equals
// v1 u1^T + v2 u2^T + (v1 x v2)(u1 x u2)^T, so only the two dominant pairs are needed.
// Returns false when sigma_2 is numerically zero (collinear points: R not unique).
// Part 1: the Jacobi sweeps.  ca[j] = column j of H V (the scaled left singular vectors), cv[j] = column j of V.
GSF_HD __forceinline__ void jacobi_sweeps(const double* H, double (&ca)[3][3], double (&cv)[3][3]) {
#pragma unroll
    for (int j = 0; j < 3; ++j)
#pragma unroll
        for (int r = 0; r < 3; ++r) { ca[j][r] = H[3 * r + j]; cv[j][r] = (r == j) ? 1.0 : 0.0; }
    // Cyclic sweeps over the pairs (0,1), (0,2), (1,2).  (A single rotation body with the columns shifted cyclically
    // through its two slots was tried for its smaller instruction footprint: the 36 register moves per rotation cost
    // more than the three inlined bodies.)
    for (int sweep = 0; sweep < 30; ++sweep) {
        bool rotated = false;
        if (jacobi_rotate_pair(ca[0], ca[1], cv[0], cv[1])) rotated = true;
        if (jacobi_rotate_pair(ca[0], ca[2], cv[0], cv[2])) rotated = true;
        if (jacobi_rotate_pair(ca[1], ca[2], cv[1], cv[2])) rotated = true;
        if (!rotated) break;
    }
}
GSF_HD __forceinline__ double det3(const double* H) {
    return H[0] * (H[4] * H[8] - H[5] * H[7]) - H[1] * (H[3] * H[8] - H[5] * H[6]) + H[2] * (H[3] * H[7] - H[4] * H[6]);
}
// Part 2: rotation and singular-value sum from the converged columns; det(H) = det(U) det(V) sigma1 sigma2 sigma3
// decides the reflection branch.
GSF_HD __forceinline__ bool rotation_from_columns(const double (&ca)[3][3], const double (&cv)[3][3], double detH, double* R,
                                                  double& sigma_sum, bool& reflected) {
    // singular values = column norms; the two dominant pairs are picked with warp-uniform branches on the column
    // registers (no index arithmetic, no select chains)
    double sg[3];
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        const double q = ca[j][0] * ca[j][0] + ca[j][1] * ca[j][1] + ca[j][2] * ca[j][2];
        sg[j] = q > 0.0 ? q * rsqrt_(q) : 0.0;
    }
    sigma_sum = sg[0] + sg[1] + sg[2];
    reflected = detH < 0.0;
    double u1[3], u2[3], v1[3], v2[3], s_hi, s_mid;
    const int imin = (sg[0] <= sg[1]) ? (sg[0] <= sg[2] ? 0 : 2) : (sg[1] <= sg[2] ? 1 : 2);
#define GSF_TAKE(a, b)                                                                                     \
    {                                                                                                      \
        const bool first = sg[a] >= sg[b];                                                                 \
        s_hi = first ? sg[a] : sg[b]; s_mid = first ? sg[b] : sg[a];                                       \
        for (int k = 0; k < 3; ++k) {                                                                      \
            u1[k] = first ? ca[a][k] : ca[b][k]; u2[k] = first ? ca[b][k] : ca[a][k];                     \
            v1[k] = first ? cv[a][k] : cv[b][k]; v2[k] = first ? cv[b][k] : cv[a][k];                     \
        }                                                                                                  \
    }
    if (imin == 0) GSF_TAKE(1, 2) else if (imin == 1) GSF_TAKE(0, 2) else GSF_TAKE(0, 1)
#undef GSF_TAKE
    bool ok = s_mid > 1e-14 * s_hi && s_hi > 0.0;
    const double r0 = s_hi > 1e-300 ? rcp_(s_hi) : 0.0, r1 = s_mid > 1e-300 ? rcp_(s_mid) : 0.0;
#pragma unroll
    for (int k = 0; k < 3; ++k) { u1[k] *= r0; u2[k] *= r1; }
    // re-orthonormalise u2 against u1 (no-op to rounding when converged)
    double dp = u1[0] * u2[0] + u1[1] * u2[1] + u1[2] * u2[2];
#pragma unroll
    for (int k = 0; k < 3; ++k) u2[k] -= dp * u1[k];
    double n2 = u2[0] * u2[0] + u2[1] * u2[1] + u2[2] * u2[2];
    if (n2 > 0.0) { double rn = rsqrt_(n2); u2[0] *= rn; u2[1] *= rn; u2[2] *= rn; }
    double u3[3] = {u1[1] * u2[2] - u1[2] * u2[1], u1[2] * u2[0] - u1[0] * u2[2], u1[0] * u2[1] - u1[1] * u2[0]};
    double v3[3] = {v1[1] * v2[2] - v1[2] * v2[1], v1[2] * v2[0] - v1[0] * v2[2], v1[0] * v2[1] - v1[1] * v2[0]};
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int c = 0; c < 3; ++c) R[3 * r + c] = v1[r] * u1[c] + v2[r] * u2[c] + v3[r] * u3[c];
    return ok;
}
GSF_HD __forceinline__ bool umeyama_rotation(const double* H, double* R, double& sigma_sum, bool& reflected) {
    double ca[3][3], cv[3][3];
    jacobi_sweeps(H, ca, cv);
    return rotation_from_columns(ca, cv, det3(H), R, sigma_sum, reflected);
}

// Finish Umeyama (EKFGPSSLAM.py:443-451) from the reduced sums.
//   n      number of points, mu_s/mu_d centroids, H centred cross-covariance (not / n),
//   ss     sum |src_c|^2.
// Outputs R (row-major), t, s; returns status bits.
GSF_HD __forceinline__ int umeyama_finish(int n, const double* mu_s, const double* mu_d, const double* H, double ss,
                                     double* R, double* t, double& s) {
    double sigma_sum; bool refl;
    bool ok = umeyama_rotation(H, R, sigma_sum, refl);
    const double var_src = ss * rcp_((double)n);
    if (var_src < 1e-12) s = 1.0;
    else { s = sigma_sum * rcp_((double)n * var_src); if (s <= 1e-6) s = 1.0; }
    double rx, ry, rz;
    mat_vec(R, mu_s[0], mu_s[1], mu_s[2], rx, ry, rz);
    t[0] = mu_d[0] - s * rx; t[1] = mu_d[1] - s * ry; t[2] = mu_d[2] - s * rz;
    return ok ? ST_OK : ST_DEGENERATE;
}

// The same from converged Jacobi columns (the fast fused kernel runs the sweeps in another warp).
GSF_HD __forceinline__ int umeyama_finish_columns(int n, const double* mu_s, const double* mu_d, const double (&ca)[3][3],
                                                  const double (&cv)[3][3], double detH, double ss, double* R, double* t, double& s) {
    double sigma_sum; bool refl;
    bool ok = rotation_from_columns(ca, cv, detH, R, sigma_sum, refl);
    const double var_src = ss * rcp_((double)n);
    if (var_src < 1e-12) s = 1.0;
    else { s = sigma_sum * rcp_((double)n * var_src); if (s <= 1e-6) s = 1.0; }
    double rx, ry, rz;
    mat_vec(R, mu_s[0], mu_s[1], mu_s[2], rx, ry, rz);
    t[0] = mu_d[0] - s * rx; t[1] = mu_d[1] - s * ry; t[2] = mu_d[2] - s * rz;
    return ok ? ST_OK : ST_DEGENERATE;
}

#ifdef __CUDACC__
// ----------------------------------------------------------------------------- reductions
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(GSF_FULL_MASK, v, o);
    return v;
}
__device__ __forceinline__ int warp_sum_i(int v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(GSF_FULL_MASK, v, o);
    return v;
}

// Deterministic block-wide sum of K doubles per thread: fixed xor-shuffle tree inside each
// warp, then warp partials combined in warp order through shared memory.  `scratch` needs
// K * (blockDim.x / 32) doubles.  Every thread returns with the totals in v[].
template <int K>
__device__ inline void block_sum(double* v, double* scratch) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
#pragma unroll
    for (int k = 0; k < K; ++k) v[k] = warp_sum(v[k]);
    if (nwarp == 1) return;
    __syncthreads();
    if (lane == 0) {
#pragma unroll
        for (int k = 0; k < K; ++k) scratch[warp * K + k] = v[k];
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < K; ++k) {
        double acc = scratch[k];
        for (int w = 1; w < nwarp; ++w) acc += scratch[w * K + k];
        v[k] = acc;
    }
}

#endif  // __CUDACC__

// Squared Euclidean distance with the rounding sequence of scipy's cdist (EKFGPSSLAM.py:1030): every square and every
// partial sum rounded on its own, (dx^2 + dy^2) + dz^2, no FMA contraction -- so the value does not depend on which of
// the products a pruning test has already computed, and the nearest-neighbour kernels agree bit for bit.
GSF_HD __forceinline__ double dist2_rn(double dx, double dy, double dz) {
#ifdef __CUDA_ARCH__
    return __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
#else
    volatile double a = dx * dx, b = dy * dy, c = dz * dz;
    volatile double ab = a + b;
    return ab + c;
#endif
}
GSF_HD __forceinline__ bool row_has_nan(double a, double b, double c) { return isnan(a) || isnan(b) || isnan(c); }

}  // namespace gsf
