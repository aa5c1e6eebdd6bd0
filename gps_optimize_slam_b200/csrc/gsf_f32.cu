// Optional fp32 mode of the fused Sim3 -> EKF path (north star: "1e-4 m in an optional fp32 mode"; SURVEY 7 H5: fp64 local
// origin, fp64 reductions).  Storage is fp32 RELATIVE to per-trajectory fp64 origins -- time since t0, SLAM positions minus
// a SLAM origin, measurements / fused positions minus a UTM origin -- so a pose costs 44 B in + 28 B out instead of 88 + 56.
// Reference path replaced (/root/reference/EKFGPSSLAM.py): Sim3 point selection :972-998 (all points; anything else is
// flagged for the fp64 path), compute_sim3_transform :428-459, row 0 of transform_trajectory :461-467,
// apply_ekf_correction :831-935 with ExtendedKalmanFilter :679-772.
//
// Organisation: ONE THREAD PER TRAJECTORY, no scans, no shared memory, no role hand-offs.  fp32 halves the bytes, and a
// lane walking its own trajectory with 128-bit loads uses every sector it touches (the other half of a sector is the same
// lane's next load, served by L1), so the time-serial recursion -- which the fp64 kernels have to turn into Moebius / affine
// scans to find parallelism inside one trajectory -- can simply run as written, 32 trajectories per warp:
//   pass 1  Umeyama sums over (pos, z), accumulated in fp64 from the fp32 inputs (exact products), per lane;
//   SVD     one-sided Jacobi per lane in fp64 (all 32 lanes busy, where the block-per-trajectory kernels keep one);
//   pass 2  the filter in fp32 in INNOVATION form e_i = x_i - z_i:  e_i = (1 - k_i)(e_{i-1} + M(C) dp_i - dz_i), so every
//           fp32 quantity is metres, not the 1e2..1e3 m of the relative coordinates; x_0 and e_0 come from fp64;
//           covariance P' = R pp / (pp + R) per axis; q_i = normalize(C (x) q^_i); residual count for the RANSAC flag.
// Trajectories that need the general machinery (a pose without GNSS -> outage / RTS, a GNSS time gap, the 180 s window,
// a time step <= 1e-6 s, fewer than min_samples points, a zero quaternion) get GSF_ST_NEEDS_FP64 and NaN outputs.
#include "gsf_common.cuh"
#include "gsf_internal.cuh"

namespace gsf {

constexpr int F32_T = 128;

// ---- format conversion: fp64 absolute -> fp32 relative (origins: t0, SLAM position and measurement of the middle pose)
__global__ void __launch_bounds__(256) to_local_f32_kernel(const double* __restrict__ ts, const double* __restrict__ pos, const double* __restrict__ quat,
                                                           const double* __restrict__ z, const long long* __restrict__ offsets, int B,
                                                           float* __restrict__ ts32, float* __restrict__ pos32, float* __restrict__ quat32,
                                                           float* __restrict__ z32, double* __restrict__ origins) {
    for (int b = blockIdx.x; b < B; b += gridDim.x) {
        const long long e0 = offsets[b];
        const int n = (int)(offsets[b + 1] - e0);
        if (n <= 0) continue;
        // origin: the first pose at or after the middle that has a measurement (else the middle pose's SLAM position only)
        __shared__ double o[7];
        if (threadIdx.x == 0) {
            int m = n / 2;
            while (m < n && row_has_nan(z[3 * (e0 + m)], z[3 * (e0 + m) + 1], z[3 * (e0 + m) + 2])) ++m;
            if (m == n) { m = n / 2; while (m > 0 && row_has_nan(z[3 * (e0 + m)], z[3 * (e0 + m) + 1], z[3 * (e0 + m) + 2])) --m; }
            const bool ok = !row_has_nan(z[3 * (e0 + m)], z[3 * (e0 + m) + 1], z[3 * (e0 + m) + 2]);
            o[0] = ts[e0];
            for (int k = 0; k < 3; ++k) { o[1 + k] = pos[3 * (e0 + m) + k]; o[4 + k] = ok ? z[3 * (e0 + m) + k] : 0.0; }
            for (int k = 0; k < 7; ++k) origins[7 * (size_t)b + k] = o[k];
        }
        __syncthreads();
        for (int i = threadIdx.x; i < n; i += blockDim.x) {
            const long long g = e0 + i;
            ts32[g] = (float)(ts[g] - o[0]);
            for (int k = 0; k < 3; ++k) { pos32[3 * g + k] = (float)(pos[3 * g + k] - o[1 + k]); z32[3 * g + k] = (float)(z[3 * g + k] - o[4 + k]); }
            for (int k = 0; k < 4; ++k) quat32[4 * g + k] = (float)quat[4 * g + k];
        }
        __syncthreads();
    }
}
__global__ void __launch_bounds__(256) from_local_f32_kernel(const float* __restrict__ p32, const long long* __restrict__ offsets, int B,
                                                             const double* __restrict__ origins, double* __restrict__ out) {
    for (int b = blockIdx.x; b < B; b += gridDim.x) {
        const long long e0 = offsets[b];
        const int n = (int)(offsets[b + 1] - e0);
        const double* o = origins + 7 * (size_t)b + 4;
        for (int i = threadIdx.x; i < 3 * n; i += blockDim.x) out[3 * e0 + i] = (double)p32[3 * e0 + i] + o[i % 3];
    }
}

struct F32Args {
    const float* ts; const float* pos; const float* quat; const float* z; const double* origins; const long long* offsets;
    const FuseParams* params; int params_per_traj; int B;
    float* out_pos; float* out_quat; double* sim3_out; int* status;
};

// 4 poses of an array with W floats per pose -> v[4 W]: 128-bit loads when the group is 16-byte aligned (trajectory offsets
// that are multiples of 4 poses), scalar loads otherwise and for the ragged tail.
template <int W>
__device__ __forceinline__ void load_group(const float* __restrict__ base, int i0, int n, bool vec, float* v) {
    if (vec && i0 + 4 <= n) {
#pragma unroll
        for (int k = 0; k < W; ++k) {
            const float4 x = __ldg(reinterpret_cast<const float4*>(base + (size_t)W * i0) + k);
            v[4 * k] = x.x; v[4 * k + 1] = x.y; v[4 * k + 2] = x.z; v[4 * k + 3] = x.w;
        }
    } else {
#pragma unroll
        for (int k = 0; k < 4 * W; ++k) v[k] = (i0 + k / W < n) ? __ldg(base + (size_t)W * i0 + k) : 0.0f;
    }
}
template <int W>
__device__ __forceinline__ void store_group(float* __restrict__ base, int i0, int n, bool vec, const float* v) {
    if (vec && i0 + 4 <= n) {
#pragma unroll
        for (int k = 0; k < W; ++k) reinterpret_cast<float4*>(base + (size_t)W * i0)[k] = make_float4(v[4 * k], v[4 * k + 1], v[4 * k + 2], v[4 * k + 3]);
    } else {
#pragma unroll
        for (int k = 0; k < 4 * W; ++k) if (i0 + k / W < n) base[(size_t)W * i0 + k] = v[k];
    }
}

__global__ void __launch_bounds__(F32_T) fuse_f32_kernel(const F32Args A) {
    const int b = blockIdx.x * F32_T + threadIdx.x;
    if (b >= A.B) return;
    const long long e0 = A.offsets[b];
    const int n = (int)(A.offsets[b + 1] - e0);
    if (n <= 0) { A.status[b] = ST_EMPTY; return; }
    const FuseParams& prm = A.params[A.params_per_traj ? b : 0];
    const bool vec = (e0 & 3) == 0;
    const float* __restrict__ ts = A.ts + e0;
    const float* __restrict__ pos = A.pos + 3 * e0;
    const float* __restrict__ z = A.z + 3 * e0;
    const float* __restrict__ quat = A.quat + 4 * e0;
    float* __restrict__ op = A.out_pos + 3 * e0;
    float* __restrict__ oq = A.out_quat + 4 * e0;
    const float NANF = __int_as_float(0x7fc00000);
    // ---- pass 1: Umeyama sums in fp64 (relative coordinates are already pivot-shifted), validity / gap / window checks
    double s_a[3] = {0, 0, 0}, s_b[3] = {0, 0, 0}, H[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0}, ss = 0.0;
    int need64 = 0;
    const float t_first = __ldg(ts);
    {
        const float gap = (float)prm.gap_threshold, tlim = t_first + (float)prm.max_duration;
        float tprev = t_first;
        float tn[4], pn[12], zn[12];
        load_group<1>(ts, 0, n, vec, tn); load_group<3>(pos, 0, n, vec, pn); load_group<3>(z, 0, n, vec, zn);
        for (int i0 = 0; i0 < n; i0 += 4) {
            float tv[4], pv[12], zv[12];
#pragma unroll
            for (int k = 0; k < 4; ++k) tv[k] = tn[k];
#pragma unroll
            for (int k = 0; k < 12; ++k) { pv[k] = pn[k]; zv[k] = zn[k]; }
            if (i0 + 4 < n) { load_group<1>(ts, i0 + 4, n, vec, tn); load_group<3>(pos, i0 + 4, n, vec, pn); load_group<3>(z, i0 + 4, n, vec, zn); }
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                if (i0 + k < n) {
                    const float t = tv[k];
                    const float z0 = zv[3 * k], z1 = zv[3 * k + 1], z2 = zv[3 * k + 2];
                    if (!(z0 == z0) || !(z1 == z1) || !(z2 == z2)) need64 = 1;
                    if (i0 + k > 0 && (!(t - tprev > 1e-6f) || t - tprev > gap)) need64 = 1;
                    if (t > tlim) need64 = 1;
                    tprev = t;
                    const double a0 = pv[3 * k], a1 = pv[3 * k + 1], a2 = pv[3 * k + 2], b0 = z0, b1 = z1, b2 = z2;
                    s_a[0] += a0; s_a[1] += a1; s_a[2] += a2; s_b[0] += b0; s_b[1] += b1; s_b[2] += b2;
                    H[0] += a0 * b0; H[1] += a0 * b1; H[2] += a0 * b2; H[3] += a1 * b0; H[4] += a1 * b1; H[5] += a1 * b2;
                    H[6] += a2 * b0; H[7] += a2 * b1; H[8] += a2 * b2;
                    ss += a0 * a0 + a1 * a1 + a2 * a2;
                }
            }
        }
    }
    const Quat q0{(double)__ldg(quat), (double)__ldg(quat + 1), (double)__ldg(quat + 2), (double)__ldg(quat + 3)};
    if (n < 3 || n < prm.min_samples || qnorm2(q0) == 0.0) need64 = 1;
    if (need64) {
        for (int i = 0; i < 3 * n; ++i) op[i] = NANF;
        for (int i = 0; i < 4 * n; ++i) oq[i] = NANF;
        A.status[b] = ST_NEEDS_FP64;
        return;
    }
    // ---- Umeyama finish in fp64 (same routine as the fp64 kernels), in the relative frames
    const double nn = (double)n, inv = 1.0 / nn;
    double mu_s[3], mu_d[3];
    for (int k = 0; k < 3; ++k) { mu_s[k] = s_a[k] * inv; mu_d[k] = s_b[k] * inv; }
    for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) H[3 * r + c] -= nn * mu_s[r] * mu_d[c];
    ss -= nn * (mu_s[0] * mu_s[0] + mu_s[1] * mu_s[1] + mu_s[2] * mu_s[2]);
    double R[9], t[3], s = 1.0;
    int st = umeyama_finish(n, mu_s, mu_d, H, ss, R, t, s);
    const Quat qR = quat_from_matrix(R);
    const Quat q0h = qunit(q0);
    const Quat qs0 = qunit_or_identity(qmul(qR, q0h));
    const Quat Cq = qmul(qs0, qconj(q0h));
    double Md[9]; qmat(Cq, Md);
    if (A.sim3_out) {
        // back to the absolute frames: t_abs = t_rel - s R o_slam + o_utm
        const double* o = A.origins + 7 * (size_t)b;
        double rx, ry, rz;
        mat_vec(R, o[1], o[2], o[3], rx, ry, rz);
        double* so = A.sim3_out + 16 * (size_t)b;
        for (int k = 0; k < 9; ++k) so[k] = R[k];
        so[9] = t[0] - s * rx + o[4]; so[10] = t[1] - s * ry + o[5]; so[11] = t[2] - s * rz + o[6];
        so[12] = s; so[13] = nn; so[14] = nn;
    }
    // ---- pass 2: the filter in fp32, innovation form
    float M[9]; for (int k = 0; k < 9; ++k) M[k] = (float)Md[k];
    const float Cx = (float)Cq.x, Cy = (float)Cq.y, Cz = (float)Cq.z, Cw = (float)Cq.w;
    const float sc = (float)s;
    float P[3] = {(float)prm.p0[0], (float)prm.p0[1], (float)prm.p0[2]};
    const float Q[3] = {(float)prm.q[0], (float)prm.q[1], (float)prm.q[2]};
    const float Rn[3] = {(float)prm.r[0], (float)prm.r[1], (float)prm.r[2]};
    const float thr2 = prm.residual_thresh > 0.0 ? (float)(prm.residual_thresh * prm.residual_thresh) : -1.0f;
    float e[3], y[3], pp[3], zp[3], tp = t_first;
    {
        // x_0 = s R p_0 + t in fp64; e_0 = x_0 - z_0 (also the Sim3 residual of pose 0)
        const double p0 = __ldg(pos), p1 = __ldg(pos + 1), p2 = __ldg(pos + 2);
        double x0d, x1d, x2d;
        mat_vec(R, p0, p1, p2, x0d, x1d, x2d);
        const double xd[3] = {s * x0d + t[0], s * x1d + t[1], s * x2d + t[2]};
#pragma unroll
        for (int a = 0; a < 3; ++a) { zp[a] = __ldg(z + a); pp[a] = __ldg(pos + a); e[a] = (float)(xd[a] - (double)zp[a]); y[a] = e[a]; }
    }
    int nviol = 0, bad = 0;
    float tn[4], pn[12], zn[12], qn[16];                    // the next group's inputs are requested before the current group is processed
    load_group<1>(ts, 0, n, vec, tn); load_group<3>(pos, 0, n, vec, pn); load_group<3>(z, 0, n, vec, zn); load_group<4>(quat, 0, n, vec, qn);
    for (int i0 = 0; i0 < n; i0 += 4) {
        float tv[4], pv[12], zv[12], qv[16], ov[12], oqv[16];
#pragma unroll
        for (int k = 0; k < 4; ++k) tv[k] = tn[k];
#pragma unroll
        for (int k = 0; k < 12; ++k) { pv[k] = pn[k]; zv[k] = zn[k]; }
#pragma unroll
        for (int k = 0; k < 16; ++k) qv[k] = qn[k];
        if (i0 + 4 < n) { load_group<1>(ts, i0 + 4, n, vec, tn); load_group<3>(pos, i0 + 4, n, vec, pn); load_group<3>(z, i0 + 4, n, vec, zn); load_group<4>(quat, i0 + 4, n, vec, qn); }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int i = i0 + k;
            if (i < n) {
                if (i > 0) {
                    const float dt = tv[k] - tp;
                    float d[3], u[3];
#pragma unroll
                    for (int a = 0; a < 3; ++a) d[a] = pv[3 * k + a] - pp[a];
#pragma unroll
                    for (int a = 0; a < 3; ++a) u[a] = M[3 * a] * d[0] + M[3 * a + 1] * d[1] + M[3 * a + 2] * d[2];
#pragma unroll
                    for (int a = 0; a < 3; ++a) {
                        const float dz = zv[3 * k + a] - zp[a];
                        y[a] += sc * u[a] - dz;                       // residual of the Sim3 image (:411)
                        const float pq = P[a] + Q[a] * dt, kk = pq / (pq + Rn[a]), om = 1.0f - kk;      // :705-731 in scalar form
                        P[a] = om * pq * om + kk * Rn[a] * kk;
                        e[a] = om * (e[a] + u[a] - dz);               // innovation form of x = om (x + u) + k z
                    }
                }
                if (thr2 > 0.0f && !(y[0] * y[0] + y[1] * y[1] + y[2] * y[2] < thr2)) ++nviol;
                tp = tv[k];
#pragma unroll
                for (int a = 0; a < 3; ++a) { pp[a] = pv[3 * k + a]; zp[a] = zv[3 * k + a]; ov[3 * k + a] = zp[a] + e[a]; }
                // q_state[i] = C (x) q^_i, normalised (telescoped odometry, see gsf_fused.cu)
                const float qx = qv[4 * k], qy = qv[4 * k + 1], qz = qv[4 * k + 2], qw = qv[4 * k + 3];
                const float n2 = qx * qx + qy * qy + qz * qz + qw * qw;
                if (n2 == 0.0f) bad = 1;
                const float rs = rsqrtf(n2);
                oqv[4 * k] = (Cw * qx + qw * Cx + Cy * qz - Cz * qy) * rs;
                oqv[4 * k + 1] = (Cw * qy + qw * Cy + Cz * qx - Cx * qz) * rs;
                oqv[4 * k + 2] = (Cw * qz + qw * Cz + Cx * qy - Cy * qx) * rs;
                oqv[4 * k + 3] = (Cw * qw - Cx * qx - Cy * qy - Cz * qz) * rs;
            }
        }
        store_group<3>(op, i0, n, vec, ov); store_group<4>(oq, i0, n, vec, oqv);
    }
    if (nviol) st |= ST_RANSAC_OUTLIERS;
    if (bad) st |= ST_BAD_QUATERNION;
    if (A.sim3_out) A.sim3_out[16 * (size_t)b + 15] = (double)nviol;
    A.status[b] = st;
}

cudaError_t launch_to_local_f32(const double* ts, const double* pos, const double* quat, const double* z, const long long* offsets, int B,
                                float* ts32, float* pos32, float* quat32, float* z32, double* origins, int num_sms, cudaStream_t stream) {
    if (B <= 0) return cudaSuccess;
    long long grid = (long long)num_sms * 8; if (grid > B) grid = B;
    to_local_f32_kernel<<<(unsigned)grid, 256, 0, stream>>>(ts, pos, quat, z, offsets, B, ts32, pos32, quat32, z32, origins);
    return cudaGetLastError();
}
cudaError_t launch_from_local_f32(const float* p32, const long long* offsets, int B, const double* origins, double* out, int num_sms, cudaStream_t stream) {
    if (B <= 0) return cudaSuccess;
    long long grid = (long long)num_sms * 8; if (grid > B) grid = B;
    from_local_f32_kernel<<<(unsigned)grid, 256, 0, stream>>>(p32, offsets, B, origins, out);
    return cudaGetLastError();
}
cudaError_t launch_fuse_f32(const float* ts, const float* pos, const float* quat, const float* z, const double* origins, const long long* offsets,
                            int B, const FuseParams* params, int params_per_traj, float* out_pos, float* out_quat, double* sim3_out, int* status,
                            cudaStream_t stream) {
    if (B <= 0) return cudaSuccess;
    F32Args a;
    a.ts = ts; a.pos = pos; a.quat = quat; a.z = z; a.origins = origins; a.offsets = offsets; a.params = params; a.params_per_traj = params_per_traj;
    a.B = B; a.out_pos = out_pos; a.out_quat = out_quat; a.sim3_out = sim3_out; a.status = status;
    cudaError_t e = cudaFuncSetAttribute(fuse_f32_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, 0);       // all of it as L1
    if (e != cudaSuccess) return e;
    fuse_f32_kernel<<<(B + F32_T - 1) / F32_T, F32_T, 0, stream>>>(a);
    return cudaGetLastError();
}

}  // namespace gsf
