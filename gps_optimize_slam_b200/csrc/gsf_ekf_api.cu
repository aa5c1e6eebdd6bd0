// Per-call surface of the reference's EKF helpers (dense 7x7 covariance, exactly as the Python objects hold it),
// batched over B independent filters / segments -- one thread each; these are API-completeness kernels behind the
// drop-in's ExtendedKalmanFilter / rts_smoother_segment / quaternion_nlerp / is_sharp_turn_in_segment, not throughput
// kernels (the batched trajectory path is gsf_fuse_batched_dev).
//
// Reference (file:line in /root/reference/EKFGPSSLAM.py):
//   ExtendedKalmanFilter._predict :702-715, ._update :717-734, blend of process_step :754-767,
//   rts_smoother_segment :777-803, quaternion_nlerp :94-105, is_sharp_turn_in_segment :808-826.
#include "gsf_common.cuh"
#include "gsf_ekf_strict.cuh"
#include "gsf_internal.cuh"

namespace gsf {

__device__ inline void sym7(double* P) {            // (P + P^T) / 2
    for (int i = 0; i < 7; ++i)
        for (int j = i + 1; j < 7; ++j) { const double m = 0.5 * (P[7 * i + j] + P[7 * j + i]); P[7 * i + j] = m; P[7 * j + i] = m; }
}
// Gauss-Jordan inverse with partial pivoting; false if singular to working precision.
template <int N>
__device__ inline bool invert(const double* A, double* Ainv) {
    double M[N][2 * N];
    for (int i = 0; i < N; ++i)
        for (int j = 0; j < N; ++j) { M[i][j] = A[N * i + j]; M[i][N + j] = (i == j) ? 1.0 : 0.0; }
    for (int c = 0; c < N; ++c) {
        int piv = c; double best = fabs(M[c][c]);
        for (int r = c + 1; r < N; ++r) if (fabs(M[r][c]) > best) { best = fabs(M[r][c]); piv = r; }
        if (!(best > 0.0)) return false;
        if (piv != c) for (int j = 0; j < 2 * N; ++j) { const double t = M[c][j]; M[c][j] = M[piv][j]; M[piv][j] = t; }
        const double inv = 1.0 / M[c][c];
        for (int j = 0; j < 2 * N; ++j) M[c][j] *= inv;
        for (int r = 0; r < N; ++r) {
            if (r == c) continue;
            const double f = M[r][c];
            if (f != 0.0) for (int j = 0; j < 2 * N; ++j) M[r][j] -= f * M[c][j];
        }
    }
    for (int i = 0; i < N; ++i)
        for (int j = 0; j < N; ++j) Ainv[N * i + j] = M[i][N + j];
    return true;
}
__device__ inline Quat normalize_quaternion(const Quat& q) {    // :697-700: identity below 1e-9
    const double n = sqrt(qnorm2(q));
    if (!(n > 1e-9)) return Quat{0.0, 0.0, 0.0, 1.0};
    return Quat{q.x / n, q.y / n, q.z / n, q.w / n};
}

// mode bit 0: predict from (state, cov); bit 1: update the predicted (or, without bit 0, the given) state with z and
// blend with weight w (w >= 1: plain update).  flags[b]: bit 0 = update applied, bit 1 = zero-norm quaternion (scipy
// raises), bit 2 = innovation covariance singular (the reference falls back to pinv: not reproduced, update skipped).
__global__ void ekf_step_kernel(int mode, const double* __restrict__ state, const double* __restrict__ cov,
                                const double* __restrict__ dp, const double* __restrict__ dq, const double* __restrict__ dt,
                                const double* __restrict__ z, const double* __restrict__ qdiag, const double* __restrict__ rdiag,
                                const double* __restrict__ w, int B, double* __restrict__ out_state, double* __restrict__ out_cov,
                                double* __restrict__ pred_state, double* __restrict__ pred_cov, int* __restrict__ flags) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    double x[7], P[49];
    for (int k = 0; k < 7; ++k) x[k] = state[7 * b + k];
    for (int k = 0; k < 49; ++k) P[k] = cov[49 * b + k];
    int fl = 0;
    if (mode & 1) {                                     // ---- _predict
        const Quat q{x[3], x[4], x[5], x[6]}, d{dq[4 * b], dq[4 * b + 1], dq[4 * b + 2], dq[4 * b + 3]};
        if (qnorm2(q) == 0.0 || qnorm2(d) == 0.0) fl |= 2;
        const Quat qu = qunit(q), du = qunit(d);
        double M[9], mx, my, mz;
        qmat(qu, M);
        mat_vec(M, dp[3 * b], dp[3 * b + 1], dp[3 * b + 2], mx, my, mz);
        x[0] += mx; x[1] += my; x[2] += mz;
        const Quat qp = normalize_quaternion(qunit(qmul(qu, du)));
        x[3] = qp.x; x[4] = qp.y; x[5] = qp.z; x[6] = qp.w;
        const double dta = fmax(fabs(dt[b]), 1e-6);
        for (int k = 0; k < 7; ++k) P[8 * k] += qdiag[k] * dta;
        sym7(P);
    }
    for (int k = 0; k < 7; ++k) pred_state[7 * b + k] = x[k];
    for (int k = 0; k < 49; ++k) pred_cov[49 * b + k] = P[k];
    const double z0 = z[3 * b], z1 = z[3 * b + 1], z2 = z[3 * b + 2];
    if ((mode & 2) && !row_has_nan(z0, z1, z2)) {       // ---- _update
        double S[9], Si[9];
        for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) S[3 * i + j] = P[7 * i + j] + (i == j ? rdiag[i] : 0.0);
        for (int i = 0; i < 3; ++i) for (int j = i + 1; j < 3; ++j) { const double m = 0.5 * (S[3 * i + j] + S[3 * j + i]); S[3 * i + j] = m; S[3 * j + i] = m; }
        if (!invert<3>(S, Si)) fl |= 4;
        else {
            double K[21];                               // P H^T S^-1: 7x3
            for (int i = 0; i < 7; ++i) for (int j = 0; j < 3; ++j) K[3 * i + j] = P[7 * i] * Si[j] + P[7 * i + 1] * Si[3 + j] + P[7 * i + 2] * Si[6 + j];
            const double inn[3] = {z0 - x[0], z1 - x[1], z2 - x[2]};
            double xu[7];
            for (int i = 0; i < 7; ++i) xu[i] = x[i] + K[3 * i] * inn[0] + K[3 * i + 1] * inn[1] + K[3 * i + 2] * inn[2];
            const Quat qn = normalize_quaternion(Quat{xu[3], xu[4], xu[5], xu[6]});
            xu[3] = qn.x; xu[4] = qn.y; xu[5] = qn.z; xu[6] = qn.w;
            // Joseph form: (I - K H) P (I - K H)^T + K R K^T
            double Am[49], T[49], Pu[49];
            for (int i = 0; i < 7; ++i) for (int j = 0; j < 7; ++j) Am[7 * i + j] = (i == j ? 1.0 : 0.0) - (j < 3 ? K[3 * i + j] : 0.0);
            for (int i = 0; i < 7; ++i) for (int j = 0; j < 7; ++j) { double a = 0.0; for (int k = 0; k < 7; ++k) a += Am[7 * i + k] * P[7 * k + j]; T[7 * i + j] = a; }
            for (int i = 0; i < 7; ++i) for (int j = 0; j < 7; ++j) {
                double a = 0.0;
                for (int k = 0; k < 7; ++k) a += T[7 * i + k] * Am[7 * j + k];
                for (int k = 0; k < 3; ++k) a += K[3 * i + k] * rdiag[k] * K[3 * j + k];
                Pu[7 * i + j] = a;
            }
            sym7(Pu);
            const double wb = w[b];
            if (wb < 1.0) {                             // blend (:754-767): lerp position, nlerp quaternion, updated covariance
                for (int k = 0; k < 3; ++k) xu[k] = (1.0 - wb) * x[k] + wb * xu[k];
                double q2[4] = {xu[3], xu[4], xu[5], xu[6]};
                const double dot = x[3] * q2[0] + x[4] * q2[1] + x[5] * q2[2] + x[6] * q2[3];
                if (dot < 0.0) for (int k = 0; k < 4; ++k) q2[k] = -q2[k];
                const double wc = fmin(fmax(wb, 0.0), 1.0);
                double qi[4];
                for (int k = 0; k < 4; ++k) qi[k] = (1.0 - wc) * x[3 + k] + wc * q2[k];
                const double nq = sqrt(qi[0] * qi[0] + qi[1] * qi[1] + qi[2] * qi[2] + qi[3] * qi[3]);
                for (int k = 0; k < 4; ++k) xu[3 + k] = nq < 1e-9 ? (wb < 0.5 ? x[3 + k] : q2[k]) : qi[k] / nq;
            }
            for (int k = 0; k < 7; ++k) x[k] = xu[k];
            for (int k = 0; k < 49; ++k) P[k] = Pu[k];
            fl |= 1;
        }
    }
    for (int k = 0; k < 7; ++k) out_state[7 * b + k] = x[k];
    for (int k = 0; k < 49; ++k) out_cov[49 * b + k] = P[k];
    flags[b] = fl;
}

// rts_smoother_segment (:777-803): one thread per segment [offsets[b], offsets[b+1]).  A singular P_pred leaves the
// filtered state in place (the reference tries pinv first).
__global__ void rts_segment_kernel(const double* __restrict__ xf, const double* __restrict__ Pf, const double* __restrict__ xp,
                                   const double* __restrict__ Pp, const long long* __restrict__ offsets, int B,
                                   double* __restrict__ xs, double* __restrict__ Ps) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    const long long e0 = offsets[b], n = offsets[b + 1] - e0;
    if (n <= 0) return;
    for (int k = 0; k < 7; ++k) xs[7 * (e0 + n - 1) + k] = xf[7 * (e0 + n - 1) + k];
    for (int k = 0; k < 49; ++k) Ps[49 * (e0 + n - 1) + k] = Pf[49 * (e0 + n - 1) + k];
    for (long long i = n - 2; i >= 0; --i) {
        const long long g = e0 + i;
        double Pi[49], Ak[49];
        const bool ok = invert<7>(Pp + 49 * (g + 1), Pi);
        if (!ok) {
            for (int k = 0; k < 7; ++k) xs[7 * g + k] = xf[7 * g + k];
            for (int k = 0; k < 49; ++k) Ps[49 * g + k] = Pf[49 * g + k];
            continue;
        }
        for (int r = 0; r < 7; ++r) for (int c = 0; c < 7; ++c) { double a = 0.0; for (int k = 0; k < 7; ++k) a += Pf[49 * g + 7 * r + k] * Pi[7 * k + c]; Ak[7 * r + c] = a; }
        double d[7], xo[7];
        for (int k = 0; k < 7; ++k) d[k] = xs[7 * (g + 1) + k] - xp[7 * (g + 1) + k];
        for (int r = 0; r < 7; ++r) { double a = xf[7 * g + r]; for (int k = 0; k < 7; ++k) a += Ak[7 * r + k] * d[k]; xo[r] = a; }
        const Quat qn = normalize_quaternion(Quat{xo[3], xo[4], xo[5], xo[6]});
        xo[3] = qn.x; xo[4] = qn.y; xo[5] = qn.z; xo[6] = qn.w;
        for (int k = 0; k < 7; ++k) xs[7 * g + k] = xo[k];
        double D[49], T[49], Po[49];
        for (int k = 0; k < 49; ++k) D[k] = Ps[49 * (g + 1) + k] - Pp[49 * (g + 1) + k];
        for (int r = 0; r < 7; ++r) for (int c = 0; c < 7; ++c) { double a = 0.0; for (int k = 0; k < 7; ++k) a += Ak[7 * r + k] * D[7 * k + c]; T[7 * r + c] = a; }
        for (int r = 0; r < 7; ++r) for (int c = 0; c < 7; ++c) { double a = Pf[49 * g + 7 * r + c]; for (int k = 0; k < 7; ++k) a += T[7 * r + k] * Ak[7 * c + k]; Po[7 * r + c] = a; }
        sym7(Po);
        for (int k = 0; k < 49; ++k) Ps[49 * g + k] = Po[k];
    }
}

__global__ void quat_nlerp_kernel(const double* __restrict__ q1, const double* __restrict__ q2, const double* __restrict__ w, long long n,
                                  double* __restrict__ out) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= n) return;
    double a[4], c[4];
    for (int k = 0; k < 4; ++k) { a[k] = q1[4 * i + k]; c[k] = q2[4 * i + k]; }
    const double dot = a[0] * c[0] + a[1] * c[1] + a[2] * c[2] + a[3] * c[3];
    if (dot < 0.0) for (int k = 0; k < 4; ++k) c[k] = -c[k];
    const double wc = fmin(fmax(w[i], 0.0), 1.0);
    double q[4];
    for (int k = 0; k < 4; ++k) q[k] = (1.0 - wc) * a[k] + wc * c[k];
    const double nq = sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
    for (int k = 0; k < 4; ++k) out[4 * i + k] = nq < 1e-9 ? (w[i] < 0.5 ? a[k] : c[k]) : q[k] / nq;
}

// is_sharp_turn_in_segment (:808-826): flag and the largest yaw rate (rad/s; the reference prints it).
__global__ void sharp_turn_kernel(const double* __restrict__ ts, const double* __restrict__ quat, const long long* __restrict__ offsets, int B,
                                  double thresh, int* __restrict__ flags, double* __restrict__ max_rate) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    const long long e0 = offsets[b], n = offsets[b + 1] - e0;
    double worst = 0.0; int invalid = 0;
    for (long long a = 1; a < n; ++a) {
        const double t1 = ts[e0 + a - 1], t2 = ts[e0 + a];
        if (t2 <= t1) continue;
        const double* p1 = quat + 4 * (e0 + a - 1); const double* p2 = quat + 4 * (e0 + a);
        const Quat q1{p1[0], p1[1], p1[2], p1[3]}, q2{p2[0], p2[1], p2[2], p2[3]};
        if (qnorm2(q1) == 0.0 || qnorm2(q2) == 0.0) { invalid = 1; break; }
        const double y1 = yaw_zyx(q1), y2 = yaw_zyx(q2);
        const double d = atan2(sin(y2 - y1), cos(y2 - y1));
        worst = fmax(worst, fabs(d / (t2 - t1)));
    }
    flags[b] = (n >= 2 && (invalid || worst > thresh)) ? 1 : 0;
    max_rate[b] = worst;
}

cudaError_t launch_ekf_step(int mode, const double* state, const double* cov, const double* dp, const double* dq, const double* dt,
                            const double* z, const double* qdiag, const double* rdiag, const double* w, int B, double* out_state,
                            double* out_cov, double* pred_state, double* pred_cov, int* flags, cudaStream_t stream) {
    if (B <= 0) return cudaSuccess;
    ekf_step_kernel<<<(B + 31) / 32, 32, 0, stream>>>(mode, state, cov, dp, dq, dt, z, qdiag, rdiag, w, B, out_state, out_cov, pred_state, pred_cov, flags);
    return cudaGetLastError();
}
cudaError_t launch_rts_segment(const double* xf, const double* Pf, const double* xp, const double* Pp, const long long* offsets, int B,
                               double* xs, double* Ps, cudaStream_t stream) {
    if (B <= 0) return cudaSuccess;
    rts_segment_kernel<<<(B + 31) / 32, 32, 0, stream>>>(xf, Pf, xp, Pp, offsets, B, xs, Ps);
    return cudaGetLastError();
}
cudaError_t launch_quat_nlerp(const double* q1, const double* q2, const double* w, long long n, double* out, cudaStream_t stream) {
    if (n <= 0) return cudaSuccess;
    quat_nlerp_kernel<<<(unsigned)((n + 127) / 128), 128, 0, stream>>>(q1, q2, w, n, out);
    return cudaGetLastError();
}
cudaError_t launch_sharp_turn(const double* ts, const double* quat, const long long* offsets, int B, double thresh, int* flags,
                              double* max_rate, cudaStream_t stream) {
    if (B <= 0) return cudaSuccess;
    sharp_turn_kernel<<<(B + 31) / 32, 32, 0, stream>>>(ts, quat, offsets, B, thresh, flags, max_rate);
    return cudaGetLastError();
}

}  // namespace gsf
