// Optional fp32 mode of the fused Sim3 -> EKF path (north star: "1e-4 m in an optional fp32 mode"; SURVEY 7 H5: fp64 local
// origin, fp64 reductions).  Storage is fp32 RELATIVE to per-trajectory fp64 origins -- time since t0, SLAM positions minus
// a SLAM origin, measurements / fused positions minus a UTM origin -- so a pose costs 44 B in + 28 B out instead of 88 + 56.
// Reference path replaced (/root/reference/EKFGPSSLAM.py): Sim3 point selection :972-998 (all points; anything else is
// flagged for the fp64 path), compute_sim3_transform :428-459, row 0 of transform_trajectory :461-467,
// apply_ekf_correction :831-935 with ExtendedKalmanFilter :679-772.
//
// Organisation: ONE THREAD PER TRAJECTORY, no scans, no shared memory, no role hand-offs.  The storage interleaves 32
// trajectories (see below), so a warp's access to one component of one pose is one 128-byte line, and the time-serial
// recursion -- which the fp64 kernels have to turn into Moebius / affine scans to find parallelism inside one trajectory --
// can simply run as written, 32 trajectories per warp:
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

// ---- storage layout: INTERLEAVED BY 32 TRAJECTORIES.  Trajectories b = 32 g + l form group g; the group is padded to the
// length L_g of its longest member and stored pose-major with the 32 members innermost:
//     ts32 [(gbase_g + i) * 32 + l],   pos32 / z32 [((gbase_g + i) * 3 + c) * 32 + l],   quat32 [((gbase_g + i) * 4 + c) * 32 + l]
// (gbase_g = sum of L over the groups before g, `group_offsets`).  A warp is a group: every load or store of one component
// of one pose is ONE 128-byte line for the whole warp.  (The first version kept the AoS layout of the fp64 path and let
// every lane walk its own trajectory with 128-bit loads: 32 sectors per warp-level load, bound by L1 wavefronts at
// 3.8e10 pose-updates/s.)  gsf_to_local_f32_dev / gsf_from_local_f32_dev convert from / to the fp64 AoS arrays.
__device__ __forceinline__ size_t il_idx(long long row, int comp, int width, int lane) { return ((size_t)row * width + comp) * 32 + lane; }

__global__ void __launch_bounds__(256) to_local_f32_kernel(const double* __restrict__ ts, const double* __restrict__ pos, const double* __restrict__ quat,
                                                           const double* __restrict__ z, const long long* __restrict__ offsets,
                                                           const long long* __restrict__ group_offsets, int B,
                                                           float* __restrict__ ts32, float* __restrict__ pos32, float* __restrict__ quat32,
                                                           float* __restrict__ z32, double* __restrict__ origins) {
    const float NANF = __int_as_float(0x7fc00000);
    for (int b = blockIdx.x; b < ((B + 31) & ~31); b += gridDim.x) {
        const int g = b >> 5, l = b & 31;
        const long long gb = group_offsets[g], Lg = group_offsets[g + 1] - gb;
        const long long e0 = b < B ? offsets[b] : 0;
        const int n = b < B ? (int)(offsets[b + 1] - e0) : 0;
        __shared__ double o[7];
        __syncthreads();
        if (threadIdx.x == 0 && n > 0) {
            // origin: the first pose at or after the middle that has a measurement (else the nearest one before it)
            int m = n / 2;
            while (m < n && row_has_nan(z[3 * (e0 + m)], z[3 * (e0 + m) + 1], z[3 * (e0 + m) + 2])) ++m;
            if (m == n) { m = n / 2; while (m > 0 && row_has_nan(z[3 * (e0 + m)], z[3 * (e0 + m) + 1], z[3 * (e0 + m) + 2])) --m; }
            const bool ok = !row_has_nan(z[3 * (e0 + m)], z[3 * (e0 + m) + 1], z[3 * (e0 + m) + 2]);
            o[0] = ts[e0];
            for (int k = 0; k < 3; ++k) { o[1 + k] = pos[3 * (e0 + m) + k]; o[4 + k] = ok ? z[3 * (e0 + m) + k] : 0.0; }
            for (int k = 0; k < 7; ++k) origins[7 * (size_t)b + k] = o[k];
        }
        __syncthreads();
        for (long long i = threadIdx.x; i < Lg; i += blockDim.x) {
            const bool in = i < n;
            const long long ge = e0 + i;
            ts32[il_idx(gb + i, 0, 1, l)] = in ? (float)(ts[ge] - o[0]) : 0.0f;
            for (int k = 0; k < 3; ++k) {
                pos32[il_idx(gb + i, k, 3, l)] = in ? (float)(pos[3 * ge + k] - o[1 + k]) : 0.0f;
                z32[il_idx(gb + i, k, 3, l)] = in ? (float)(z[3 * ge + k] - o[4 + k]) : NANF;
            }
            for (int k = 0; k < 4; ++k) quat32[il_idx(gb + i, k, 4, l)] = in ? (float)quat[4 * ge + k] : 0.0f;
        }
    }
}
__global__ void __launch_bounds__(256) from_local_f32_kernel(const float* __restrict__ p32, const float* __restrict__ q32,
                                                             const long long* __restrict__ offsets, const long long* __restrict__ group_offsets,
                                                             int B, const double* __restrict__ origins, double* __restrict__ out_pos, double* __restrict__ out_quat) {
    for (int b = blockIdx.x; b < B; b += gridDim.x) {
        const int g = b >> 5, l = b & 31;
        const long long gb = group_offsets[g];
        const long long e0 = offsets[b];
        const int n = (int)(offsets[b + 1] - e0);
        const double* o = origins + 7 * (size_t)b + 4;
        for (int i = threadIdx.x; i < n; i += blockDim.x) {
            if (out_pos) for (int k = 0; k < 3; ++k) out_pos[3 * (e0 + i) + k] = (double)p32[il_idx(gb + i, k, 3, l)] + o[k];
            if (out_quat && q32) for (int k = 0; k < 4; ++k) out_quat[4 * (e0 + i) + k] = (double)q32[il_idx(gb + i, k, 4, l)];
        }
    }
}

struct F32Args {
    const float* ts; const float* pos; const float* quat; const float* z; const double* origins; const long long* offsets;
    const long long* group_offsets;
    const FuseParams* params; int params_per_traj; int B;
    float* out_pos; float* out_quat; double* sim3_out; int* status;
};

constexpr int F32_G = 4;          // poses requested together per lane (44 coalesced loads in flight)

__global__ void __launch_bounds__(F32_T) fuse_f32_kernel(const F32Args A) {
    const int b = blockIdx.x * F32_T + threadIdx.x;
    const int lane = threadIdx.x & 31;
    const int g = b >> 5;
    if (g * 32 >= A.B) return;
    const bool live = b < A.B;
    const long long gb = A.group_offsets[g];
    const int Lg = (int)(A.group_offsets[g + 1] - gb);       // padded length of the group (warp-uniform loop bound)
    const long long e0 = live ? A.offsets[b] : 0;
    const int n = live ? (int)(A.offsets[b + 1] - e0) : 0;
    const FuseParams& prm = A.params[(A.params_per_traj && live) ? b : 0];
    const float* __restrict__ ts = A.ts; const float* __restrict__ pos = A.pos; const float* __restrict__ z = A.z; const float* __restrict__ quat = A.quat;
    float* __restrict__ op = A.out_pos; float* __restrict__ oq = A.out_quat;
    const float NANF = __int_as_float(0x7fc00000);
    // ---- pass 1: Umeyama sums in fp64 (relative coordinates are already pivot-shifted), validity / gap / window checks
    double s_a[3] = {0, 0, 0}, s_b[3] = {0, 0, 0}, H[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0}, ss = 0.0;
    int need64 = 0;
    const float t_first = Lg > 0 ? __ldg(ts + il_idx(gb, 0, 1, lane)) : 0.0f;
    {
        const float gap = (float)prm.gap_threshold, tlim = t_first + (float)prm.max_duration;
        float tprev = t_first;
        for (int i0 = 0; i0 < Lg; i0 += F32_G) {
            float tv[F32_G], pv[F32_G][3], zv[F32_G][3];
#pragma unroll
            for (int k = 0; k < F32_G; ++k) {
                const long long row = gb + min(i0 + k, Lg - 1);
                tv[k] = __ldg(ts + il_idx(row, 0, 1, lane));
#pragma unroll
                for (int c = 0; c < 3; ++c) { pv[k][c] = __ldg(pos + il_idx(row, c, 3, lane)); zv[k][c] = __ldg(z + il_idx(row, c, 3, lane)); }
            }
#pragma unroll
            for (int k = 0; k < F32_G; ++k) {
                if (i0 + k < n) {
                    const float t = tv[k];
                    const float z0 = zv[k][0], z1 = zv[k][1], z2 = zv[k][2];
                    if (!(z0 == z0) || !(z1 == z1) || !(z2 == z2)) need64 = 1;
                    if (i0 + k > 0 && (!(t - tprev > 1e-6f) || t - tprev > gap)) need64 = 1;
                    if (t > tlim) need64 = 1;
                    tprev = t;
                    const double a0 = pv[k][0], a1 = pv[k][1], a2 = pv[k][2], b0 = z0, b1 = z1, b2 = z2;
                    s_a[0] += a0; s_a[1] += a1; s_a[2] += a2; s_b[0] += b0; s_b[1] += b1; s_b[2] += b2;
                    H[0] += a0 * b0; H[1] += a0 * b1; H[2] += a0 * b2; H[3] += a1 * b0; H[4] += a1 * b1; H[5] += a1 * b2;
                    H[6] += a2 * b0; H[7] += a2 * b1; H[8] += a2 * b2;
                    ss += a0 * a0 + a1 * a1 + a2 * a2;
                }
            }
        }
    }
    Quat q0{0.0, 0.0, 0.0, 1.0};
    if (n > 0) q0 = Quat{(double)__ldg(quat + il_idx(gb, 0, 4, lane)), (double)__ldg(quat + il_idx(gb, 1, 4, lane)),
                         (double)__ldg(quat + il_idx(gb, 2, 4, lane)), (double)__ldg(quat + il_idx(gb, 3, 4, lane))};
    if (n < 3 || n < prm.min_samples || qnorm2(q0) == 0.0) need64 = 1;
    // ---- Umeyama finish in fp64 (same routine as the fp64 kernels), in the relative frames; lanes that cannot use it idle
    double R[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1}, t[3] = {0, 0, 0}, s = 1.0;
    int st = 0;
    Quat Cq{0.0, 0.0, 0.0, 1.0}, qs0{0.0, 0.0, 0.0, 1.0};
    double Md[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
    if (!need64) {
        const double nn = (double)n, inv = 1.0 / nn;
        double mu_s[3], mu_d[3];
        for (int k = 0; k < 3; ++k) { mu_s[k] = s_a[k] * inv; mu_d[k] = s_b[k] * inv; }
        for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) H[3 * r + c] -= nn * mu_s[r] * mu_d[c];
        ss -= nn * (mu_s[0] * mu_s[0] + mu_s[1] * mu_s[1] + mu_s[2] * mu_s[2]);
        st = umeyama_finish(n, mu_s, mu_d, H, ss, R, t, s);
        const Quat qR = quat_from_matrix(R);
        const Quat q0h = qunit(q0);
        qs0 = qunit_or_identity(qmul(qR, q0h));
        Cq = qmul(qs0, qconj(q0h));
        qmat(Cq, Md);
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
    }
    // ---- pass 2: the filter in fp32, innovation form
    float M[9]; for (int k = 0; k < 9; ++k) M[k] = (float)Md[k];
    const float Cx = (float)Cq.x, Cy = (float)Cq.y, Cz = (float)Cq.z, Cw = (float)Cq.w;
    const float sc = (float)s;
    float P[3] = {(float)prm.p0[0], (float)prm.p0[1], (float)prm.p0[2]};
    const float Q[3] = {(float)prm.q[0], (float)prm.q[1], (float)prm.q[2]};
    const float Rn[3] = {(float)prm.r[0], (float)prm.r[1], (float)prm.r[2]};
    const float thr2 = prm.residual_thresh > 0.0 ? (float)(prm.residual_thresh * prm.residual_thresh) : -1.0f;
    float e[3] = {0, 0, 0}, y[3] = {0, 0, 0}, pp[3] = {0, 0, 0}, zp[3] = {0, 0, 0}, tp = t_first;
    if (!need64) {
        // x_0 = s R p_0 + t in fp64; e_0 = x_0 - z_0 (also the Sim3 residual of pose 0)
        double p0d[3];
        for (int a = 0; a < 3; ++a) { pp[a] = __ldg(pos + il_idx(gb, a, 3, lane)); zp[a] = __ldg(z + il_idx(gb, a, 3, lane)); p0d[a] = pp[a]; }
        double x0d, x1d, x2d;
        mat_vec(R, p0d[0], p0d[1], p0d[2], x0d, x1d, x2d);
        const double xd[3] = {s * x0d + t[0], s * x1d + t[1], s * x2d + t[2]};
#pragma unroll
        for (int a = 0; a < 3; ++a) { e[a] = (float)(xd[a] - (double)zp[a]); y[a] = e[a]; }
    }
    int nviol = 0, bad = 0;
    for (int i0 = 0; i0 < Lg; i0 += F32_G) {
        float tv[F32_G], pv[F32_G][3], zv[F32_G][3], qv[F32_G][4];
#pragma unroll
        for (int k = 0; k < F32_G; ++k) {
            const long long row = gb + min(i0 + k, Lg - 1);
            tv[k] = __ldg(ts + il_idx(row, 0, 1, lane));
#pragma unroll
            for (int c = 0; c < 3; ++c) { pv[k][c] = __ldg(pos + il_idx(row, c, 3, lane)); zv[k][c] = __ldg(z + il_idx(row, c, 3, lane)); }
#pragma unroll
            for (int c = 0; c < 4; ++c) qv[k][c] = __ldg(quat + il_idx(row, c, 4, lane));
        }
#pragma unroll
        for (int k = 0; k < F32_G; ++k) {
            const int i = i0 + k;
            if (i >= Lg) break;
            float ox[3] = {NANF, NANF, NANF}, oqv[4] = {NANF, NANF, NANF, NANF};
            if (i < n && !need64) {
                if (i > 0) {
                    const float dt = tv[k] - tp;
                    float d[3], u[3];
#pragma unroll
                    for (int a = 0; a < 3; ++a) d[a] = pv[k][a] - pp[a];
#pragma unroll
                    for (int a = 0; a < 3; ++a) u[a] = M[3 * a] * d[0] + M[3 * a + 1] * d[1] + M[3 * a + 2] * d[2];
#pragma unroll
                    for (int a = 0; a < 3; ++a) {
                        const float dz = zv[k][a] - zp[a];
                        y[a] += sc * u[a] - dz;                       // residual of the Sim3 image (:411)
                        const float pq = P[a] + Q[a] * dt, kk = pq / (pq + Rn[a]), om = 1.0f - kk;      // :705-731 in scalar form
                        P[a] = om * pq * om + kk * Rn[a] * kk;
                        e[a] = om * (e[a] + u[a] - dz);               // innovation form of x = om (x + u) + k z
                    }
                }
                if (thr2 > 0.0f && !(y[0] * y[0] + y[1] * y[1] + y[2] * y[2] < thr2)) ++nviol;
                tp = tv[k];
#pragma unroll
                for (int a = 0; a < 3; ++a) { pp[a] = pv[k][a]; zp[a] = zv[k][a]; ox[a] = zp[a] + e[a]; }
                // q_state[i] = C (x) q^_i, normalised (telescoped odometry, see gsf_fused.cu)
                const float qx = qv[k][0], qy = qv[k][1], qz = qv[k][2], qw = qv[k][3];
                const float n2 = qx * qx + qy * qy + qz * qz + qw * qw;
                if (n2 == 0.0f) bad = 1;
                const float rs = rsqrtf(n2);
                oqv[0] = (Cw * qx + qw * Cx + Cy * qz - Cz * qy) * rs;
                oqv[1] = (Cw * qy + qw * Cy + Cz * qx - Cx * qz) * rs;
                oqv[2] = (Cw * qz + qw * Cz + Cx * qy - Cy * qx) * rs;
                oqv[3] = (Cw * qw - Cx * qx - Cy * qy - Cz * qz) * rs;
            }
            const long long row = gb + i;                             // full 128-byte lines: padding rows and refused trajectories get NaN
#pragma unroll
            for (int a = 0; a < 3; ++a) op[il_idx(row, a, 3, lane)] = ox[a];
#pragma unroll
            for (int a = 0; a < 4; ++a) oq[il_idx(row, a, 4, lane)] = oqv[a];
        }
    }
    if (!live) return;
    if (n <= 0) { A.status[b] = ST_EMPTY; return; }
    if (need64) { A.status[b] = ST_NEEDS_FP64; return; }
    if (nviol) st |= ST_RANSAC_OUTLIERS;
    if (bad) st |= ST_BAD_QUATERNION;
    if (A.sim3_out) A.sim3_out[16 * (size_t)b + 15] = (double)nviol;
    A.status[b] = st;
}

cudaError_t launch_to_local_f32(const double* ts, const double* pos, const double* quat, const double* z, const long long* offsets,
                                const long long* group_offsets, int B, float* ts32, float* pos32, float* quat32, float* z32, double* origins,
                                int num_sms, cudaStream_t stream) {
    if (B <= 0) return cudaSuccess;
    long long grid = (long long)num_sms * 8; if (grid > ((B + 31) & ~31)) grid = (B + 31) & ~31;
    to_local_f32_kernel<<<(unsigned)grid, 256, 0, stream>>>(ts, pos, quat, z, offsets, group_offsets, B, ts32, pos32, quat32, z32, origins);
    return cudaGetLastError();
}
cudaError_t launch_from_local_f32(const float* p32, const float* q32, const long long* offsets, const long long* group_offsets, int B,
                                  const double* origins, double* out_pos, double* out_quat, int num_sms, cudaStream_t stream) {
    if (B <= 0) return cudaSuccess;
    long long grid = (long long)num_sms * 8; if (grid > B) grid = B;
    from_local_f32_kernel<<<(unsigned)grid, 256, 0, stream>>>(p32, q32, offsets, group_offsets, B, origins, out_pos, out_quat);
    return cudaGetLastError();
}
cudaError_t launch_fuse_f32(const float* ts, const float* pos, const float* quat, const float* z, const double* origins, const long long* offsets,
                            const long long* group_offsets, int B, const FuseParams* params, int params_per_traj, float* out_pos, float* out_quat,
                            double* sim3_out, int* status, cudaStream_t stream) {
    if (B <= 0) return cudaSuccess;
    F32Args a;
    a.ts = ts; a.pos = pos; a.quat = quat; a.z = z; a.origins = origins; a.offsets = offsets; a.group_offsets = group_offsets;
    a.params = params; a.params_per_traj = params_per_traj;
    a.B = B; a.out_pos = out_pos; a.out_quat = out_quat; a.sim3_out = sim3_out; a.status = status;
    const int Bp = (B + 31) & ~31;
    fuse_f32_kernel<<<(Bp + F32_T - 1) / F32_T, F32_T, 0, stream>>>(a);
    return cudaGetLastError();
}

}  // namespace gsf
