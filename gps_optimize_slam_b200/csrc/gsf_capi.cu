// C ABI of libgsf.so (declared in include/gsf.h): argument checks, launch configuration and
// the host-buffer pipeline.  No torch types, no hidden device allocation in the *_dev calls.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include <mutex>
#include <algorithm>

#include "../../include/gsf.h"
#include "gsf_common.cuh"
#include "gsf_internal.cuh"


static_assert(sizeof(gsf_fuse_params) == sizeof(gsf::FuseParams), "ABI struct mismatch");

namespace {

thread_local std::string g_err;
int fail(int code, const std::string& msg) { g_err = msg; return code; }
int cuda_fail(cudaError_t e, const char* where) {
    return fail(GSF_E_CUDA, std::string(where) + ": " + cudaGetErrorString(e));
}

struct DeviceInfo { int ok = 0; int sms = 0; int max_smem = 0; };
DeviceInfo& device_info() {
    static thread_local DeviceInfo cached[64];
    static thread_local DeviceInfo none;
    int dev = -1;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) { cudaGetLastError(); return none; }
    DeviceInfo& d = cached[dev];
    if (!d.ok) {
        cudaDeviceProp p;
        if (cudaGetDeviceProperties(&p, dev) != cudaSuccess) { cudaGetLastError(); return none; }
        if (p.major != 10) return none;                 // sm_100a cubin only: no fallback path
        d.sms = p.multiProcessorCount;
        d.max_smem = (int)p.sharedMemPerBlockOptin;
        d.ok = 1;
    }
    return d;
}
long long* g_phase_clock = nullptr;     // debug hook, see gsf_debug_phase_clock
bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// WGS84 constants of the Krueger series (same numbers as oracle/utm_kruger.py).
gsf::UtmConst utm_const(int zone, int south) {
    const double a = 6378137.0, f = 1.0 / 298.257223563, k0 = 0.9996;
    const double n = f / (2.0 - f), n2 = n * n, n3 = n2 * n, n4 = n3 * n, n5 = n4 * n, n6 = n5 * n;
    gsf::UtmConst K;
    K.A_k0 = k0 * a / (1.0 + n) * (1.0 + n2 / 4.0 + n4 / 64.0 + n6 / 256.0);
    K.e2 = f * (2.0 - f); K.e = sqrt(K.e2);
    K.alpha[0] = n / 2 - 2 * n2 / 3 + 5 * n3 / 16 + 41 * n4 / 180 - 127 * n5 / 288 + 7891 * n6 / 37800;
    K.alpha[1] = 13 * n2 / 48 - 3 * n3 / 5 + 557 * n4 / 1440 + 281 * n5 / 630 - 1983433 * n6 / 1935360;
    K.alpha[2] = 61 * n3 / 240 - 103 * n4 / 140 + 15061 * n5 / 26880 + 167603 * n6 / 181440;
    K.alpha[3] = 49561 * n4 / 161280 - 179 * n5 / 168 + 6601661 * n6 / 7257600;
    K.alpha[4] = 34729 * n5 / 80640 - 3418889 * n6 / 1995840;
    K.alpha[5] = 212378941 * n6 / 319334400;
    K.beta[0] = n / 2 - 2 * n2 / 3 + 37 * n3 / 96 - n4 / 360 - 81 * n5 / 512 + 96199 * n6 / 604800;
    K.beta[1] = n2 / 48 + n3 / 15 - 437 * n4 / 1440 + 46 * n5 / 105 - 1118711 * n6 / 3870720;
    K.beta[2] = 17 * n3 / 480 - 37 * n4 / 840 - 209 * n5 / 4480 + 5569 * n6 / 90720;
    K.beta[3] = 4397 * n4 / 161280 - 11 * n5 / 504 - 830251 * n6 / 7257600;
    K.beta[4] = 4583 * n5 / 161280 - 108847 * n6 / 3991680;
    K.beta[5] = 20648693 * n6 / 638668800;
    K.lon0 = (6.0 * zone - 183.0) * 0.017453292519943295769236907684886;
    K.fn = south ? 10000000.0 : 0.0;
    return K;
}

// Threads per trajectory for the fused kernel: enough blocks per SM to keep ~16 warps
// resident given the shared-memory footprint of one trajectory.
int pick_threads(int cap, int max_smem) {
    // Every thread owns an odd-length chunk of poses; scans and reductions cost per thread, so
    // use the fewest threads that still give ~12 resident warps per SM for the shared-memory
    // footprint of one trajectory (chunk length stays around 9 poses).
    const size_t smem = gsf::fuse_smem_bytes(cap);
    int blocks = (int)((size_t)max_smem / (smem + 1024));
    if (blocks < 1) blocks = 1;
    int threads = 32;
    while (threads < 256 && (blocks * threads < 384 || threads * 9 < cap)) threads <<= 1;
    const char* force = getenv("GSF_FUSE_THREADS");            // tuning hook
    if (force && atoi(force) > 0) threads = atoi(force);
    return threads;
}

}  // namespace

extern "C" {

const char* gsf_version(void) { return "gsf 0.1 (sm_100a)"; }
const char* gsf_last_error(void) { return g_err.c_str(); }
int gsf_device_sm_count(void) {
    DeviceInfo& d = device_info();
    if (!d.ok) return fail(GSF_E_NO_DEVICE, "no sm_100 CUDA device (libgsf has no CPU fallback)");
    return d.sms;
}

int gsf_fuse_batched_dev(const double* ts, const double* pos, const double* quat, const double* z,
                         const int64_t* offsets, int32_t B, int64_t max_len,
                         const gsf_fuse_params* params, int32_t params_per_traj,
                         const double* init_pos, const double* init_quat,
                         double* out_pos, double* out_quat, double* sim3_out, int32_t* status, void* stream) {
    DeviceInfo& d = device_info();
    if (!d.ok) return fail(GSF_E_NO_DEVICE, "no sm_100 CUDA device (libgsf has no CPU fallback)");
    if (B == 0) return 0;
    if (B < 0 || max_len < 0 || !ts || !pos || !quat || !z || !offsets || !params || !out_pos || !out_quat || !status)
        return fail(GSF_E_INVALID, "gsf_fuse_batched_dev: null pointer or negative size");
    if ((init_pos == nullptr) != (init_quat == nullptr))
        return fail(GSF_E_INVALID, "gsf_fuse_batched_dev: init_pos and init_quat must be given together");
    if (!aligned16(quat) || !aligned16(out_quat))
        return fail(GSF_E_INVALID, "gsf_fuse_batched_dev: quaternion arrays must be 16-byte aligned");
    // Trajectories that fit the shared-memory staging (about 4000 poses) run on the shared-memory kernels; longer ones
    // are left to the tiled kernel launched behind them (any length, like the reference: EKFGPSSLAM.py:831-935).
    int cap = (int)std::max<int64_t>(std::min<int64_t>(max_len, 1 << 20), 2);
    bool need_long = max_len > cap;
    if (gsf::fuse_smem_bytes(cap) > (size_t)d.max_smem) {
        need_long = true;
        cap = (int)(((size_t)d.max_smem - 3400) / 57) & ~1;
        while (cap > 2 && gsf::fuse_smem_bytes(cap) > (size_t)d.max_smem) cap -= 2;
    }
    gsf::FuseArgs a;
    a.ts = ts; a.pos = pos; a.quat = quat; a.z = z;
    a.offsets = reinterpret_cast<const long long*>(offsets);
    a.params = reinterpret_cast<const gsf::FuseParams*>(params); a.params_per_traj = params_per_traj;
    a.init_pos = init_pos; a.init_quat = init_quat;
    a.out_pos = out_pos; a.out_quat = out_quat; a.sim3_out = sim3_out; a.status = status;
    a.B = B; a.cap = cap;
    a.use_tma = aligned16(ts) && aligned16(pos) && aligned16(z) && aligned16(out_pos);
    a.phase_clock = g_phase_clock;
    a.only_deferred = 0; a.defer_count = nullptr; a.work_counter = nullptr;
    // Fast kernel first (all-valid trajectories); whatever it defers goes through the general kernel
    // behind it on the same stream.  GSF_FUSE_IMPL=general forces the general kernel for everything.
    const char* impl = getenv("GSF_FUSE_IMPL");
    const bool want_fast = !(impl && strcmp(impl, "general") == 0);
    cudaError_t e;
    int counter_slot = -1;
    if (want_fast && a.use_tma && !init_pos && gsf::fast_fuse_supported(cap, d.max_smem)) {
        e = gsf::defer_counter(&a.defer_count, &counter_slot);
        if (e == cudaErrorNotReady) { a.defer_count = nullptr; counter_slot = -1; }      // every counter slot busy: general kernel for the whole batch
        else if (e != cudaSuccess) return cuda_fail(e, "gsf_fuse_batched_dev (defer counter)");
        else {
            a.work_counter = a.defer_count + 1;
            e = cudaMemsetAsync(a.defer_count, 0, 4 * sizeof(int), (cudaStream_t)stream);
            if (e != cudaSuccess) return cuda_fail(e, "gsf_fuse_batched_dev (defer counter reset)");
            e = gsf::launch_fuse_fast(a, d.sms, (cudaStream_t)stream);
            if (e != cudaSuccess) return cuda_fail(e, "gsf_fuse_batched_dev (fast kernel)");
            a.only_deferred = 1;
            a.work_counter = a.defer_count + 2;           // the deferred pass hands out chunks of trajectories from its own counter
        }
    }
    e = gsf::launch_fuse(a, pick_threads(cap, d.max_smem), d.sms, (cudaStream_t)stream);
    if (e != cudaSuccess) return cuda_fail(e, "gsf_fuse_batched_dev");
    if (counter_slot >= 0) gsf::defer_counter_release(counter_slot, (cudaStream_t)stream);
    if (need_long) {
        e = gsf::launch_fuse_long(a, d.sms, (cudaStream_t)stream);
        if (e != cudaSuccess) return cuda_fail(e, "gsf_fuse_batched_dev (long-trajectory kernel)");
    }
    return 0;
}

int64_t gsf_hypothesis_grid_work_doubles(int64_t n, int32_t H) {
    return gsf::grid_work_doubles(n < 0 ? 0 : n, H < 0 ? 0 : H);
}
int gsf_ekf_hypothesis_grid_dev(const double* ts, const double* pos, const double* quat, const double* z, int64_t n,
                                const gsf_fuse_params* params, int32_t H, double* work, double* stats,
                                double* sim3_out, int32_t* status, void* stream) {
    DeviceInfo& d = device_info();
    if (!d.ok) return fail(GSF_E_NO_DEVICE, "no sm_100 CUDA device (libgsf has no CPU fallback)");
    if (n < 2 || H <= 0 || !ts || !pos || !quat || !z || !params || !work || !stats || !status)
        return fail(GSF_E_INVALID, "gsf_ekf_hypothesis_grid_dev: null pointer or bad size");
    if (!aligned16(work)) return fail(GSF_E_INVALID, "gsf_ekf_hypothesis_grid_dev: work must be 16-byte aligned");
    cudaError_t e = gsf::launch_hypothesis_grid(ts, pos, quat, z, n, reinterpret_cast<const gsf::FuseParams*>(params), H, work, stats,
                                                sim3_out, status, d.max_smem, d.sms, (cudaStream_t)stream);
    if (e == cudaErrorInvalidValue) { cudaGetLastError(); return fail(GSF_E_TOO_LARGE, "gsf_ekf_hypothesis_grid_dev: trajectory too long for the shared-memory candidate set"); }
    if (e != cudaSuccess) return cuda_fail(e, "gsf_ekf_hypothesis_grid_dev");
    return 0;
}

int64_t gsf_noise_grid_work_doubles(int64_t n, int32_t Kq, int32_t Kz, int32_t Kr, int64_t h_first, int64_t h_count) {
    if (n < 0 || Kq <= 0 || Kz <= 0 || Kr <= 0 || h_first < 0 || h_count <= 0 || h_first + h_count > (int64_t)Kq * Kz * Kr) return 0;
    return gsf::noise_grid_work_doubles(n, Kq, Kz, Kr, h_first, h_count);
}
int gsf_ekf_noise_grid_dev(const double* ts, const double* pos, const double* quat, const double* z, int64_t n,
                           const gsf_fuse_params* base, const double* q_xy, const double* q_z, const double* r,
                           int32_t Kq, int32_t Kz, int32_t Kr, int64_t h_first, int64_t h_count,
                           double* work, double* stats, double* sim3_out, int32_t* status, void* stream) {
    DeviceInfo& d = device_info();
    if (!d.ok) return fail(GSF_E_NO_DEVICE, "no sm_100 CUDA device (libgsf has no CPU fallback)");
    if (n < 2 || Kq <= 0 || Kz <= 0 || Kr <= 0 || h_first < 0 || h_count <= 0 || h_first + h_count > (int64_t)Kq * Kz * Kr || !ts || !pos ||
        !quat || !z || !base || !q_xy || !q_z || !r || !work || !stats || !status)
        return fail(GSF_E_INVALID, "gsf_ekf_noise_grid_dev: null pointer or bad size");
    if (!aligned16(work)) return fail(GSF_E_INVALID, "gsf_ekf_noise_grid_dev: work must be 16-byte aligned");
    cudaError_t e = gsf::launch_noise_grid(ts, pos, quat, z, n, reinterpret_cast<const gsf::FuseParams*>(base), q_xy, q_z, r, Kq, Kz, Kr,
                                           h_first, h_count, work, stats, sim3_out, status, d.max_smem, d.sms, (cudaStream_t)stream);
    if (e == cudaErrorInvalidValue) { cudaGetLastError(); return fail(GSF_E_TOO_LARGE, "gsf_ekf_noise_grid_dev: trajectory too long for the shared-memory candidate set"); }
    if (e != cudaSuccess) return cuda_fail(e, "gsf_ekf_noise_grid_dev");
    return 0;
}

int gsf_poly_ransac_dev(const double* t, const double* y, int32_t y_stride, const int32_t* window_idx, const int64_t* fit_offsets,
                        const int32_t* fit_axis, const int32_t* samples, const int32_t* dyn_trials, const int64_t* dyn_offsets,
                        int32_t F, int32_t min_samples, int32_t degree, int32_t max_trials, double residual_threshold,
                        uint8_t* inlier_mask, int32_t* n_trials, int32_t* status, void* stream) {
    DeviceInfo& d = device_info();
    if (!d.ok) return fail(GSF_E_NO_DEVICE, "no sm_100 CUDA device (libgsf has no CPU fallback)");
    if (F == 0) return 0;
    if (F < 0 || !t || !y || y_stride < 1 || !window_idx || !fit_offsets || !fit_axis || !samples || !dyn_trials || !dyn_offsets ||
        !inlier_mask || !n_trials || !status || max_trials < 1 || degree < 0 || degree > 3 || min_samples < degree + 1)
        return fail(GSF_E_INVALID, "gsf_poly_ransac_dev: null pointer or bad size (degree <= 3, min_samples > degree)");
    cudaError_t e = gsf::launch_poly_ransac(t, y, y_stride, window_idx, reinterpret_cast<const long long*>(fit_offsets), fit_axis, samples,
                                            dyn_trials, reinterpret_cast<const long long*>(dyn_offsets), F, min_samples, degree, max_trials,
                                            residual_threshold, inlier_mask, n_trials, status, (cudaStream_t)stream);
    if (e != cudaSuccess) return cuda_fail(e, "gsf_poly_ransac_dev");
    return 0;
}

int64_t gsf_parse_table_work_bytes(int64_t nbytes) { return gsf::parse_table_work_bytes(nbytes < 0 ? 0 : nbytes); }
int gsf_parse_table_dev(const char* text, int64_t nbytes, int32_t delimiter, int32_t max_cols, double* out, int64_t max_rows,
                        void* work, int64_t* info, void* stream) {
    DeviceInfo& d = device_info();
    if (!d.ok) return fail(GSF_E_NO_DEVICE, "no sm_100 CUDA device (libgsf has no CPU fallback)");
    if (nbytes < 0 || max_cols < 1 || max_rows < 0 || (nbytes > 0 && !text) || !out || !work || !info || delimiter < 0 || delimiter > 127)
        return fail(GSF_E_INVALID, "gsf_parse_table_dev: null pointer or bad size");
    cudaError_t e = gsf::launch_parse_table(text, nbytes, delimiter, max_cols, out, max_rows, work, reinterpret_cast<long long*>(info), (cudaStream_t)stream);
    if (e != cudaSuccess) return cuda_fail(e, "gsf_parse_table_dev");
    return 0;
}
int64_t gsf_write_rows_work_bytes(int64_t n) { return gsf::write_rows_work_bytes(n < 0 ? 0 : n); }
int gsf_write_pose_rows_dev(const double* ts, const double* xyz, const double* quat, int64_t n, const int32_t* decimals,
                            const char* header, int32_t header_bytes, char* out, int64_t capacity, void* work, int64_t* out_info,
                            void* stream) {
    DeviceInfo& d = device_info();
    if (!d.ok) return fail(GSF_E_NO_DEVICE, "no sm_100 CUDA device (libgsf has no CPU fallback)");
    if (n < 0 || !ts || !xyz || !quat || !decimals || !out || !work || !out_info || header_bytes < 0 || (header_bytes > 0 && !header))
        return fail(GSF_E_INVALID, "gsf_write_pose_rows_dev: null pointer or bad size");
    cudaError_t e = gsf::launch_write_pose_rows(ts, xyz, quat, n, decimals, header, header_bytes, out, capacity, work,
                                                reinterpret_cast<long long*>(out_info), (cudaStream_t)stream);
    if (e == cudaErrorInvalidValue) { cudaGetLastError(); return fail(GSF_E_INVALID, "gsf_write_pose_rows_dev: decimals must be 0..9 and the header must fit"); }
    if (e != cudaSuccess) return cuda_fail(e, "gsf_write_pose_rows_dev");
    return 0;
}

int gsf_to_local_f32_dev(const double* ts, const double* pos, const double* quat, const double* z, const int64_t* offsets,
                         const int64_t* group_offsets, int32_t B, float* ts32, float* pos32, float* quat32, float* z32, double* origins,
                         void* stream) {
    DeviceInfo& d = device_info();
    if (!d.ok) return fail(GSF_E_NO_DEVICE, "no sm_100 CUDA device (libgsf has no CPU fallback)");
    if (B == 0) return 0;
    if (B < 0 || !ts || !pos || !quat || !z || !offsets || !group_offsets || !ts32 || !pos32 || !quat32 || !z32 || !origins)
        return fail(GSF_E_INVALID, "gsf_to_local_f32_dev: null pointer or negative size");
    cudaError_t e = gsf::launch_to_local_f32(ts, pos, quat, z, reinterpret_cast<const long long*>(offsets),
                                             reinterpret_cast<const long long*>(group_offsets), B, ts32, pos32, quat32, z32, origins,
                                             d.sms, (cudaStream_t)stream);
    if (e != cudaSuccess) return cuda_fail(e, "gsf_to_local_f32_dev");
    return 0;
}
int gsf_from_local_f32_dev(const float* pos32, const float* quat32, const int64_t* offsets, const int64_t* group_offsets, int32_t B,
                           const double* origins, double* pos, double* quat, void* stream) {
    DeviceInfo& d = device_info();
    if (!d.ok) return fail(GSF_E_NO_DEVICE, "no sm_100 CUDA device (libgsf has no CPU fallback)");
    if (B == 0) return 0;
    if (B < 0 || !pos32 || !offsets || !group_offsets || !origins || (!pos && !quat) || (quat && !quat32))
        return fail(GSF_E_INVALID, "gsf_from_local_f32_dev: null pointer or negative size");
    cudaError_t e = gsf::launch_from_local_f32(pos32, quat32, reinterpret_cast<const long long*>(offsets),
                                               reinterpret_cast<const long long*>(group_offsets), B, origins, pos, quat, d.sms, (cudaStream_t)stream);
    if (e != cudaSuccess) return cuda_fail(e, "gsf_from_local_f32_dev");
    return 0;
}
int gsf_fuse_batched_f32_dev(const float* ts32, const float* pos32, const float* quat32, const float* z32, const double* origins,
                             const int64_t* offsets, const int64_t* group_offsets, int32_t B, const gsf_fuse_params* params,
                             int32_t params_per_traj, float* out_pos32, float* out_quat32, double* sim3_out, int32_t* status, void* stream) {
    DeviceInfo& d = device_info();
    if (!d.ok) return fail(GSF_E_NO_DEVICE, "no sm_100 CUDA device (libgsf has no CPU fallback)");
    if (B == 0) return 0;
    if (B < 0 || !ts32 || !pos32 || !quat32 || !z32 || !origins || !offsets || !group_offsets || !params || !out_pos32 || !out_quat32 || !status)
        return fail(GSF_E_INVALID, "gsf_fuse_batched_f32_dev: null pointer or negative size");
    cudaError_t e = gsf::launch_fuse_f32(ts32, pos32, quat32, z32, origins, reinterpret_cast<const long long*>(offsets),
                                         reinterpret_cast<const long long*>(group_offsets), B,
                                         reinterpret_cast<const gsf::FuseParams*>(params), params_per_traj, out_pos32, out_quat32, sim3_out,
                                         status, (cudaStream_t)stream);
    if (e != cudaSuccess) return cuda_fail(e, "gsf_fuse_batched_f32_dev");
    return 0;
}

int gsf_ekf_strict_batched_dev(const double* ts, const double* pos, const double* quat, const double* z,
                               const int64_t* offsets, int32_t B,
                               const gsf_fuse_params* params, int32_t params_per_traj,
                               const double* init_pos, const double* init_quat,
                               double* out_pos, double* out_quat, int32_t* status, void* stream) {
    DeviceInfo& d = device_info();
    if (!d.ok) return fail(GSF_E_NO_DEVICE, "no sm_100 CUDA device (libgsf has no CPU fallback)");
    if (B == 0) return 0;
    if (B < 0 || !ts || !pos || !quat || !z || !offsets || !params || !init_pos || !init_quat || !out_pos || !out_quat || !status)
        return fail(GSF_E_INVALID, "gsf_ekf_strict_batched_dev: null pointer or negative size");
    cudaError_t e = gsf::launch_ekf_strict(ts, pos, quat, z, reinterpret_cast<const long long*>(offsets),
                                           reinterpret_cast<const gsf::FuseParams*>(params), params_per_traj,
                                           init_pos, init_quat, out_pos, out_quat, status, B, (cudaStream_t)stream);
    if (e != cudaSuccess) return cuda_fail(e, "gsf_ekf_strict_batched_dev");
    return 0;
}

int gsf_ekf_step_dev(int32_t mode, const double* state, const double* cov, const double* motion_dp, const double* motion_dq,
                     const double* dt, const double* z, const double* q_diag_per_sec, const double* r_diag,
                     const double* blend_w, int32_t B, double* out_state, double* out_cov, double* pred_state,
                     double* pred_cov, int32_t* flags, void* stream) {
    DeviceInfo& d = device_info();
    if (!d.ok) return fail(GSF_E_NO_DEVICE, "no sm_100 CUDA device (libgsf has no CPU fallback)");
    if (B == 0) return 0;
    if (B < 0 || !state || !cov || !motion_dp || !motion_dq || !dt || !z || !q_diag_per_sec || !r_diag || !blend_w || !out_state ||
        !out_cov || !pred_state || !pred_cov || !flags || !(mode & 3))
        return fail(GSF_E_INVALID, "gsf_ekf_step_dev: null pointer, negative size or empty mode");
    cudaError_t e = gsf::launch_ekf_step(mode, state, cov, motion_dp, motion_dq, dt, z, q_diag_per_sec, r_diag, blend_w, B, out_state,
                                         out_cov, pred_state, pred_cov, flags, (cudaStream_t)stream);
    if (e != cudaSuccess) return cuda_fail(e, "gsf_ekf_step_dev");
    return 0;
}
int gsf_rts_segment_dev(const double* x_filt, const double* P_filt, const double* x_pred, const double* P_pred,
                        const int64_t* offsets, int32_t B, double* x_smooth, double* P_smooth, void* stream) {
    DeviceInfo& d = device_info();
    if (!d.ok) return fail(GSF_E_NO_DEVICE, "no sm_100 CUDA device (libgsf has no CPU fallback)");
    if (B == 0) return 0;
    if (B < 0 || !x_filt || !P_filt || !x_pred || !P_pred || !offsets || !x_smooth || !P_smooth)
        return fail(GSF_E_INVALID, "gsf_rts_segment_dev: null pointer or negative size");
    cudaError_t e = gsf::launch_rts_segment(x_filt, P_filt, x_pred, P_pred, reinterpret_cast<const long long*>(offsets), B, x_smooth,
                                            P_smooth, (cudaStream_t)stream);
    if (e != cudaSuccess) return cuda_fail(e, "gsf_rts_segment_dev");
    return 0;
}
int gsf_quat_nlerp_dev(const double* q1, const double* q2, const double* weight_q2, int64_t n, double* out, void* stream) {
    DeviceInfo& d = device_info();
    if (!d.ok) return fail(GSF_E_NO_DEVICE, "no sm_100 CUDA device (libgsf has no CPU fallback)");
    if (n == 0) return 0;
    if (n < 0 || !q1 || !q2 || !weight_q2 || !out) return fail(GSF_E_INVALID, "gsf_quat_nlerp_dev: null pointer or negative size");
    cudaError_t e = gsf::launch_quat_nlerp(q1, q2, weight_q2, n, out, (cudaStream_t)stream);
    if (e != cudaSuccess) return cuda_fail(e, "gsf_quat_nlerp_dev");
    return 0;
}
int gsf_sharp_turn_dev(const double* ts, const double* quat, const int64_t* offsets, int32_t B, double yaw_rate_threshold,
                       int32_t* flags, double* max_rate, void* stream) {
    DeviceInfo& d = device_info();
    if (!d.ok) return fail(GSF_E_NO_DEVICE, "no sm_100 CUDA device (libgsf has no CPU fallback)");
    if (B == 0) return 0;
    if (B < 0 || !ts || !quat || !offsets || !flags || !max_rate) return fail(GSF_E_INVALID, "gsf_sharp_turn_dev: null pointer or negative size");
    cudaError_t e = gsf::launch_sharp_turn(ts, quat, reinterpret_cast<const long long*>(offsets), B, yaw_rate_threshold, flags, max_rate,
                                           (cudaStream_t)stream);
    if (e != cudaSuccess) return cuda_fail(e, "gsf_sharp_turn_dev");
    return 0;
}

int64_t gsf_umeyama_work_doubles(int32_t B, int64_t max_len) {
    return (int64_t)B * gsf::sim3_tiles_for(max_len) * 20;
}

int gsf_sim3_umeyama_batched_dev(const double* src, const double* dst, const int64_t* offsets,
                                 const uint8_t* mask, int32_t B, int64_t max_len, double* work,
                                 double* R, double* t, double* s, int32_t* status, void* stream) {
    DeviceInfo& d = device_info();
    if (!d.ok) return fail(GSF_E_NO_DEVICE, "no sm_100 CUDA device (libgsf has no CPU fallback)");
    if (B == 0) return 0;
    if (B < 0 || max_len < 0 || !src || !dst || !offsets || !work || !R || !t || !s || !status)
        return fail(GSF_E_INVALID, "gsf_sim3_umeyama_batched_dev: null pointer or negative size");
    cudaError_t e = gsf::launch_umeyama(src, dst, reinterpret_cast<const long long*>(offsets), mask, B, max_len, work,
                                        R, t, s, status, (cudaStream_t)stream);
    if (e != cudaSuccess) return cuda_fail(e, "gsf_sim3_umeyama_batched_dev");
    return 0;
}

int gsf_sim3_partial_stats_dev(const double* src, const double* dst, const int64_t* offsets2, const uint8_t* mask, int64_t n,
                               double* work, double* stats17, void* stream) {
    DeviceInfo& d = device_info();
    if (!d.ok) return fail(GSF_E_NO_DEVICE, "no sm_100 CUDA device (libgsf has no CPU fallback)");
    if (n < 0 || !src || !dst || !offsets2 || !work || !stats17) return fail(GSF_E_INVALID, "gsf_sim3_partial_stats_dev: null pointer or negative size");
    cudaError_t e = gsf::launch_sim3_partial_stats(src, dst, reinterpret_cast<const long long*>(offsets2), mask, n, work, stats17, (cudaStream_t)stream);
    if (e != cudaSuccess) return cuda_fail(e, "gsf_sim3_partial_stats_dev");
    return 0;
}
int gsf_sim3_from_partial_stats_dev(const double* stats, int32_t shards, double* R, double* t, double* s, int32_t* status, void* stream) {
    DeviceInfo& d = device_info();
    if (!d.ok) return fail(GSF_E_NO_DEVICE, "no sm_100 CUDA device (libgsf has no CPU fallback)");
    if (shards <= 0 || !stats || !R || !t || !s || !status) return fail(GSF_E_INVALID, "gsf_sim3_from_partial_stats_dev: null pointer or bad shard count");
    cudaError_t e = gsf::launch_sim3_from_partial_stats(stats, shards, R, t, s, status, (cudaStream_t)stream);
    if (e != cudaSuccess) return cuda_fail(e, "gsf_sim3_from_partial_stats_dev");
    return 0;
}

int64_t gsf_sim3_ransac_work_doubles(int32_t trials, int64_t n) {
    return gsf::sim3_ransac_work_doubles(trials < 0 ? 0 : trials, n < 0 ? 0 : n);
}
int gsf_sim3_ransac_dev(const double* src, const double* dst, int64_t n, const int32_t* samples,
                        int32_t trials, int32_t min_samples, double residual_threshold, int32_t min_inliers,
                        double* work, uint8_t* inlier_mask, double* R, double* t, double* s,
                        int32_t* info, int32_t* status, void* stream) {
    DeviceInfo& d = device_info();
    if (!d.ok) return fail(GSF_E_NO_DEVICE, "no sm_100 CUDA device (libgsf has no CPU fallback)");
    if (n <= 0 || trials <= 0 || min_samples <= 0 || !src || !dst || !samples || !work || !inlier_mask || !R || !t || !s || !info || !status)
        return fail(GSF_E_INVALID, "gsf_sim3_ransac_dev: null pointer or non-positive size");
    cudaError_t e = gsf::launch_sim3_ransac(src, dst, n, samples, min_samples, trials, residual_threshold, min_inliers, work,
                                            inlier_mask, R, t, s, info, status, d.sms, (cudaStream_t)stream);
    if (e != cudaSuccess) return cuda_fail(e, "gsf_sim3_ransac_dev");
    return 0;
}

int gsf_sim3_apply_dev(const double* pos, const double* quat, const int64_t* offsets,
                       const double* R, const double* t, const double* s, int32_t B, int64_t max_len,
                       double* out_pos, double* out_quat, int32_t* status, void* stream) {
    DeviceInfo& d = device_info();
    if (!d.ok) return fail(GSF_E_NO_DEVICE, "no sm_100 CUDA device (libgsf has no CPU fallback)");
    if (B == 0) return 0;
    if (B < 0 || !pos || !quat || !offsets || !R || !t || !s || !out_pos || !out_quat || !status)
        return fail(GSF_E_INVALID, "gsf_sim3_apply_dev: null pointer or negative size");
    if (!aligned16(quat) || !aligned16(out_quat))
        return fail(GSF_E_INVALID, "gsf_sim3_apply_dev: quaternion arrays must be 16-byte aligned");
    cudaError_t e = gsf::launch_sim3_apply(pos, quat, reinterpret_cast<const long long*>(offsets), R, t, s, B, max_len,
                                           out_pos, out_quat, status, d.sms, (cudaStream_t)stream);
    if (e != cudaSuccess) return cuda_fail(e, "gsf_sim3_apply_dev");
    return 0;
}

int64_t gsf_ate_work_doubles(int64_t total_poses, int64_t max_len) {
    DeviceInfo& d = device_info();
    if (!d.ok) return 0;
    return max_len > gsf::ate_smem_capacity(d.max_smem) ? 4 * std::max<int64_t>(total_poses, 0) : 0;
}
int gsf_ate_nn_batched_dev(const double* traj, const double* cand, const double* ts,
                           const int64_t* offsets, int32_t B, int64_t max_len, double skip,
                           double* work, double* stats, void* stream) {
    DeviceInfo& d = device_info();
    if (!d.ok) return fail(GSF_E_NO_DEVICE, "no sm_100 CUDA device (libgsf has no CPU fallback)");
    if (B == 0) return 0;
    if (B < 0 || max_len < 0 || !traj || !cand || !ts || !offsets || !stats)
        return fail(GSF_E_INVALID, "gsf_ate_nn_batched_dev: null pointer or negative size");
    gsf::AteArgs a;
    a.traj = traj; a.cand = cand; a.ts = ts; a.offsets = reinterpret_cast<const long long*>(offsets);
    a.skip = skip; a.stats = stats; a.B = B;
    const bool big = max_len > gsf::ate_smem_capacity(d.max_smem);
    if (big && !work)
        return fail(GSF_E_INVALID, "gsf_ate_nn_batched_dev: trajectories of " + std::to_string(max_len) +
                                   " poses need the workspace of gsf_ate_work_doubles()");
    a.cap = big ? 0 : (int)std::max<int64_t>(max_len, 32);
    a.work = big ? work : nullptr;
    cudaError_t e = gsf::launch_ate(a, d.sms, (cudaStream_t)stream);
    if (e != cudaSuccess) return cuda_fail(e, "gsf_ate_nn_batched_dev");
    return 0;
}

int gsf_utm_forward_dev(const double* lon, const double* lat, int64_t n, int32_t zone, int32_t south,
                        double* east, double* north, void* stream) {
    DeviceInfo& d = device_info();
    if (!d.ok) return fail(GSF_E_NO_DEVICE, "no sm_100 CUDA device (libgsf has no CPU fallback)");
    if (n == 0) return 0;
    if (n < 0 || !lon || !lat || !east || !north || zone < 1 || zone > 60)
        return fail(GSF_E_INVALID, "gsf_utm_forward_dev: bad argument");
    cudaError_t e = gsf::launch_utm(false, lon, lat, n, utm_const(zone, south), east, north, d.sms, (cudaStream_t)stream);
    if (e != cudaSuccess) return cuda_fail(e, "gsf_utm_forward_dev");
    return 0;
}
int gsf_utm_inverse_dev(const double* east, const double* north, int64_t n, int32_t zone, int32_t south,
                        double* lon, double* lat, void* stream) {
    DeviceInfo& d = device_info();
    if (!d.ok) return fail(GSF_E_NO_DEVICE, "no sm_100 CUDA device (libgsf has no CPU fallback)");
    if (n == 0) return 0;
    if (n < 0 || !lon || !lat || !east || !north || zone < 1 || zone > 60)
        return fail(GSF_E_INVALID, "gsf_utm_inverse_dev: bad argument");
    cudaError_t e = gsf::launch_utm(true, east, north, n, utm_const(zone, south), lon, lat, d.sms, (cudaStream_t)stream);
    if (e != cudaSuccess) return cuda_fail(e, "gsf_utm_inverse_dev");
    return 0;
}
int gsf_gnss_rows_to_utm_dev(const double* rows, int64_t n, double* part, double* zone_out,
                             double* out_ts, double* out_xyz, void* stream) {
    DeviceInfo& d = device_info();
    if (!d.ok) return fail(GSF_E_NO_DEVICE, "no sm_100 CUDA device (libgsf has no CPU fallback)");
    if (n <= 0 || !rows || !part || !zone_out || !out_xyz) return fail(GSF_E_INVALID, "gsf_gnss_rows_to_utm_dev: bad argument");
    if (!aligned16(rows)) return fail(GSF_E_INVALID, "gsf_gnss_rows_to_utm_dev: rows must be 16-byte aligned");
    int nparts = (int)std::min<int64_t>(GSF_GEO_PARTS, (n + 255) / 256);
    cudaError_t e = gsf::launch_gnss_rows(rows, n, utm_const(31, 0), part, nparts, zone_out, out_ts, out_xyz, d.sms, (cudaStream_t)stream);
    if (e != cudaSuccess) return cuda_fail(e, "gsf_gnss_rows_to_utm_dev");
    return 0;
}
int gsf_geo_zone_dev(const double* lon, const double* lat, int64_t n, double* part, double* out, void* stream) {
    DeviceInfo& d = device_info();
    if (!d.ok) return fail(GSF_E_NO_DEVICE, "no sm_100 CUDA device (libgsf has no CPU fallback)");
    if (n <= 0 || !lon || !lat || !part || !out) return fail(GSF_E_INVALID, "gsf_geo_zone_dev: bad argument");
    int nparts = (int)std::min<int64_t>(GSF_GEO_PARTS, (n + 255) / 256);
    cudaError_t e = gsf::launch_geo_mean(lon, lat, n, part, nparts, out, (cudaStream_t)stream);
    if (e != cudaSuccess) return cuda_fail(e, "gsf_geo_zone_dev");
    return 0;
}

int gsf_associate_spline_dev(const double* gps_t, const double* gps_xyz, const int64_t* gps_offsets,
                             const double* slam_t, const int64_t* slam_offsets, int32_t B, double gap,
                             double* work, double* aligned, uint8_t* valid, void* stream) {
    DeviceInfo& d = device_info();
    if (!d.ok) return fail(GSF_E_NO_DEVICE, "no sm_100 CUDA device (libgsf has no CPU fallback)");
    if (B == 0) return 0;
    if (B < 0 || !gps_t || !gps_xyz || !gps_offsets || !slam_t || !slam_offsets || !work || !aligned || !valid)
        return fail(GSF_E_INVALID, "gsf_associate_spline_dev: null pointer or negative size");
    gsf::AssocArgs a;
    a.gps_t = gps_t; a.gps_xyz = gps_xyz; a.gps_off = reinterpret_cast<const long long*>(gps_offsets);
    a.slam_t = slam_t; a.slam_off = reinterpret_cast<const long long*>(slam_offsets);
    a.gap = gap; a.aligned = aligned; a.valid = valid; a.work = work; a.B = B;
    cudaError_t e = gsf::launch_associate(a, d.sms, (cudaStream_t)stream);
    if (e != cudaSuccess) return cuda_fail(e, "gsf_associate_spline_dev");
    return 0;
}

int64_t gsf_associate_spline_long_work_doubles(int64_t M, int64_t N) { return (M < 0 || N < 0) ? -1 : gsf::associate_long_work_doubles(M, N); }

int gsf_associate_spline_long_dev(const double* gps_t, const double* gps_xyz, int64_t M, const double* slam_t, int64_t N, double gap,
                                  double* work, double* aligned, uint8_t* valid, int32_t* status, void* stream) {
    DeviceInfo& d = device_info();
    if (!d.ok) return fail(GSF_E_NO_DEVICE, "no sm_100 CUDA device (libgsf has no CPU fallback)");
    if (M < 0 || N < 0 || (M > 0 && (!gps_t || !gps_xyz)) || (N > 0 && (!slam_t || !aligned || !valid)) || !work)
        return fail(GSF_E_INVALID, "gsf_associate_spline_long_dev: null pointer or negative size");
    cudaError_t e = gsf::launch_associate_long(gps_t, gps_xyz, M, slam_t, N, gap, work, aligned, valid, status, (cudaStream_t)stream);
    if (e != cudaSuccess) return cuda_fail(e, "gsf_associate_spline_long_dev");
    return 0;
}

int gsf_synth_generate_dev(double* ts, double* pos, double* quat, double* z, int64_t first_traj,
                           int32_t B, int32_t n, double dt, double speed, uint64_t seed,
                           double outage_prob, int32_t outage_max_len, void* stream) {
    DeviceInfo& d = device_info();
    if (!d.ok) return fail(GSF_E_NO_DEVICE, "no sm_100 CUDA device (libgsf has no CPU fallback)");
    if (B == 0) return 0;
    if (B < 0 || n <= 0 || !ts || !pos || !quat || !z) return fail(GSF_E_INVALID, "gsf_synth_generate_dev: bad argument");
    gsf::SynthArgs a;
    a.ts = ts; a.pos = pos; a.quat = quat; a.z = z; a.traj0 = first_traj; a.B = B; a.n = n; a.dt = dt; a.speed = speed;
    a.seed = seed; a.outage_prob = outage_prob; a.outage_max_len = outage_max_len;
    cudaError_t e = gsf::launch_synth(a, (cudaStream_t)stream);
    if (e != cudaSuccess) return cuda_fail(e, "gsf_synth_generate_dev");
    return 0;
}

// Debug hook (not part of include/gsf.h): device buffer of 16 int64 receiving clock64() stamps
// of block 0 at the phase boundaries of one trajectory; NULL switches it off.
void gsf_debug_phase_clock(long long* dev_buf) { g_phase_clock = dev_buf; }

// ----------------------------------------------------------------------------- host-buffer pipeline
namespace {
struct Slot {
    cudaStream_t stream = nullptr;
    double *ts = nullptr, *pos = nullptr, *quat = nullptr, *z = nullptr, *opos = nullptr, *oquat = nullptr;
    double *sim3 = nullptr, *ipos = nullptr, *iquat = nullptr;
    long long* off = nullptr; int* status = nullptr; gsf::FuseParams* params = nullptr;
    long long* off_host = nullptr;              // pinned
    size_t pose_cap = 0, traj_cap = 0;
};
Slot g_slots[2];
int g_slot_device = -1;
std::mutex g_slot_mutex;          // the workspace is process-global: concurrent callers of the host entry are serialised

void slot_free(Slot& s) {
    cudaFree(s.ts); cudaFree(s.pos); cudaFree(s.quat); cudaFree(s.z); cudaFree(s.opos); cudaFree(s.oquat);
    cudaFree(s.sim3); cudaFree(s.ipos); cudaFree(s.iquat); cudaFree(s.off); cudaFree(s.status); cudaFree(s.params);
    if (s.off_host) cudaFreeHost(s.off_host);
    if (s.stream) cudaStreamDestroy(s.stream);
    s = Slot();
}
cudaError_t slot_reserve(Slot& s, size_t poses, size_t trajs) {
    cudaError_t e;
    if (!s.stream && (e = cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking)) != cudaSuccess) return e;
    if (poses > s.pose_cap) {
        cudaFree(s.ts); cudaFree(s.pos); cudaFree(s.quat); cudaFree(s.z); cudaFree(s.opos); cudaFree(s.oquat);
        s.pose_cap = 0;
        const size_t p = poses + 16;
        if ((e = cudaMalloc(&s.ts, p * 8)) != cudaSuccess) return e;
        if ((e = cudaMalloc(&s.pos, p * 24)) != cudaSuccess) return e;
        if ((e = cudaMalloc(&s.quat, p * 32)) != cudaSuccess) return e;
        if ((e = cudaMalloc(&s.z, p * 24)) != cudaSuccess) return e;
        if ((e = cudaMalloc(&s.opos, p * 24)) != cudaSuccess) return e;
        if ((e = cudaMalloc(&s.oquat, p * 32)) != cudaSuccess) return e;
        s.pose_cap = poses;
    }
    if (trajs > s.traj_cap) {
        cudaFree(s.sim3); cudaFree(s.ipos); cudaFree(s.iquat); cudaFree(s.off); cudaFree(s.status); cudaFree(s.params);
        if (s.off_host) cudaFreeHost(s.off_host);
        s.traj_cap = 0; s.off_host = nullptr;
        if ((e = cudaMalloc(&s.sim3, trajs * 16 * 8)) != cudaSuccess) return e;
        if ((e = cudaMalloc(&s.ipos, trajs * 24)) != cudaSuccess) return e;
        if ((e = cudaMalloc(&s.iquat, trajs * 32)) != cudaSuccess) return e;
        if ((e = cudaMalloc(&s.off, (trajs + 1) * 8)) != cudaSuccess) return e;
        if ((e = cudaMalloc(&s.status, trajs * 4)) != cudaSuccess) return e;
        if ((e = cudaMalloc(&s.params, trajs * sizeof(gsf::FuseParams))) != cudaSuccess) return e;
        if ((e = cudaMallocHost(&s.off_host, (trajs + 1) * 8)) != cudaSuccess) return e;
        s.traj_cap = trajs;
    }
    return cudaSuccess;
}
}  // namespace

static void host_workspace_free_locked() { slot_free(g_slots[0]); slot_free(g_slots[1]); g_slot_device = -1; }
void gsf_host_workspace_free(void) {
    std::lock_guard<std::mutex> lock(g_slot_mutex);
    host_workspace_free_locked();
}

int gsf_fuse_batched_host(const double* ts, const double* pos, const double* quat, const double* z,
                          const int64_t* offsets, int32_t B, int64_t max_len,
                          const gsf_fuse_params* params, int32_t params_per_traj,
                          const double* init_pos, const double* init_quat,
                          double* out_pos, double* out_quat, double* sim3_out, int32_t* status) {
    DeviceInfo& d = device_info();
    if (!d.ok) return fail(GSF_E_NO_DEVICE, "no sm_100 CUDA device (libgsf has no CPU fallback)");
    if (B == 0) return 0;
    if (B < 0 || max_len < 0 || !ts || !pos || !quat || !z || !offsets || !params || !out_pos || !out_quat || !status)
        return fail(GSF_E_INVALID, "gsf_fuse_batched_host: null pointer or negative size");
    if ((init_pos == nullptr) != (init_quat == nullptr))
        return fail(GSF_E_INVALID, "gsf_fuse_batched_host: init_pos and init_quat must be given together");
    std::lock_guard<std::mutex> lock(g_slot_mutex);
    int dev = 0; cudaGetDevice(&dev);
    if (g_slot_device != dev) { host_workspace_free_locked(); g_slot_device = dev; }

    // chunk the batch: ~2M poses (176 MB in / 112 MB out) per chunk, two chunks in flight
    const int64_t CHUNK_POSES = 2 << 20;
    std::vector<int32_t> cuts; cuts.push_back(0);
    {
        int32_t b = 0;
        while (b < B) {
            int32_t e = b; int64_t poses = 0;
            while (e < B && (e == b || poses + (offsets[e + 1] - offsets[e]) <= CHUNK_POSES)) { poses += offsets[e + 1] - offsets[e]; ++e; }
            cuts.push_back(e); b = e;
        }
    }
    size_t max_poses = 0, max_trajs = 0;
    for (size_t c = 0; c + 1 < cuts.size(); ++c) {
        max_poses = std::max<size_t>(max_poses, (size_t)(offsets[cuts[c + 1]] - offsets[cuts[c]]));
        max_trajs = std::max<size_t>(max_trajs, (size_t)(cuts[c + 1] - cuts[c]));
    }
    for (int k = 0; k < 2; ++k) {
        cudaError_t e = slot_reserve(g_slots[k], max_poses, max_trajs);
        if (e != cudaSuccess) { host_workspace_free_locked(); return cuda_fail(e, "gsf_fuse_batched_host(workspace)"); }
    }
    int rc = 0;
    for (size_t c = 0; c + 1 < cuts.size() && rc == 0; ++c) {
        Slot& s = g_slots[c & 1];
        cudaError_t e = cudaStreamSynchronize(s.stream);            // slot's previous chunk fully drained
        if (e != cudaSuccess) { rc = cuda_fail(e, "gsf_fuse_batched_host(sync)"); break; }
        const int32_t b0 = cuts[c], nb = cuts[c + 1] - cuts[c];
        const int64_t p0 = offsets[b0], np = offsets[b0 + nb] - p0;
        for (int32_t i = 0; i <= nb; ++i) s.off_host[i] = offsets[b0 + i] - p0;
        cudaError_t ce = cudaSuccess;
        auto copy = [&](void* dst, const void* src, size_t bytes, cudaMemcpyKind kind) {
            if (ce == cudaSuccess) ce = cudaMemcpyAsync(dst, src, bytes, kind, s.stream);
        };
        copy(s.off, s.off_host, (size_t)(nb + 1) * 8, cudaMemcpyHostToDevice);
        copy(s.ts, ts + p0, (size_t)np * 8, cudaMemcpyHostToDevice);
        copy(s.pos, pos + 3 * p0, (size_t)np * 24, cudaMemcpyHostToDevice);
        copy(s.quat, quat + 4 * p0, (size_t)np * 32, cudaMemcpyHostToDevice);
        copy(s.z, z + 3 * p0, (size_t)np * 24, cudaMemcpyHostToDevice);
        copy(s.params, params + (params_per_traj ? b0 : 0), (params_per_traj ? (size_t)nb : 1) * sizeof(gsf_fuse_params), cudaMemcpyHostToDevice);
        if (init_pos) {
            copy(s.ipos, init_pos + 3 * (size_t)b0, (size_t)nb * 24, cudaMemcpyHostToDevice);
            copy(s.iquat, init_quat + 4 * (size_t)b0, (size_t)nb * 32, cudaMemcpyHostToDevice);
        }
        if (ce != cudaSuccess) { rc = cuda_fail(ce, "gsf_fuse_batched_host(H2D copy)"); break; }
        rc = gsf_fuse_batched_dev(s.ts, s.pos, s.quat, s.z, reinterpret_cast<const int64_t*>(s.off), nb, max_len,
                                  reinterpret_cast<const gsf_fuse_params*>(s.params), params_per_traj,
                                  init_pos ? s.ipos : nullptr, init_pos ? s.iquat : nullptr,
                                  s.opos, s.oquat, s.sim3, s.status, s.stream);
        if (rc != 0) break;
        copy(out_pos + 3 * p0, s.opos, (size_t)np * 24, cudaMemcpyDeviceToHost);
        copy(out_quat + 4 * p0, s.oquat, (size_t)np * 32, cudaMemcpyDeviceToHost);
        if (sim3_out) copy(sim3_out + 16 * (size_t)b0, s.sim3, (size_t)nb * 128, cudaMemcpyDeviceToHost);
        copy(status + b0, s.status, (size_t)nb * 4, cudaMemcpyDeviceToHost);
        if (ce != cudaSuccess) rc = cuda_fail(ce, "gsf_fuse_batched_host(D2H copy)");
    }
    for (int k = 0; k < 2; ++k) {
        if (!g_slots[k].stream) continue;
        cudaError_t e = cudaStreamSynchronize(g_slots[k].stream);
        if (e != cudaSuccess && rc == 0) rc = cuda_fail(e, "gsf_fuse_batched_host(final sync)");
    }
    return rc;
}

}  // extern "C"
