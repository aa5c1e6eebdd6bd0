// Launch-argument records and launcher prototypes shared by the kernel translation units and
// the C-ABI layer (gsf_capi.cu).
#pragma once
#include <cuda_runtime.h>
#include "gsf_common.cuh"

namespace gsf {
struct FuseArgs {
    const double* ts; const double* pos; const double* quat; const double* z;
    const long long* offsets;
    const FuseParams* params; int params_per_traj;
    const double* init_pos; const double* init_quat;
    double* out_pos; double* out_quat;
    double* sim3_out;
    int* status;
    int B; int cap;
    int use_tma;
    long long* phase_clock;   // debug: per-phase clock64() of block 0 (NULL = off)
    int only_deferred;        // general kernel: process only trajectories whose status is ST_DEFERRED
    int* defer_count;         // fast kernel: += deferred trajectories; general kernel (only_deferred): exit when 0
    int* work_counter;        // fast kernel: next trajectory index to hand out (zeroed before the launch)
};
size_t fuse_smem_bytes(int cap);
cudaError_t launch_fuse(const FuseArgs& a, int threads, int num_sms, cudaStream_t stream);
// Tiled kernel for trajectories longer than a.cap (gsf_long.cu): any length.
cudaError_t launch_fuse_long(const FuseArgs& a, int num_sms, cudaStream_t stream);
// Warp-specialised fast kernel (gsf_fast.cu); fast_fuse_supported: an instantiation covers `cap`.
bool fast_fuse_supported(int cap, int max_smem);
cudaError_t launch_fuse_fast(const FuseArgs& a, int num_sms, cudaStream_t stream);
cudaError_t defer_counter(int** out, int* slot_out);
void defer_counter_release(int slot, cudaStream_t stream);
cudaError_t launch_ekf_strict(const double*, const double*, const double*, const double*, const long long*,
                              const FuseParams*, int, const double*, const double*, double*, double*, int*, int, cudaStream_t);
long long grid_work_doubles(long long n, int H);
cudaError_t launch_hypothesis_grid(const double*, const double*, const double*, const double*, long long, const FuseParams*, int,
                                   double*, double*, double*, int*, int, int, cudaStream_t);
long long noise_grid_work_doubles(long long n, int Kq, int Kz, int Kr, long long h_first, long long h_count);
cudaError_t launch_noise_grid(const double*, const double*, const double*, const double*, long long, const FuseParams*, const double*,
                              const double*, const double*, int, int, int, long long, long long, double*, double*, double*, int*, int, int,
                              cudaStream_t);
cudaError_t launch_ekf_step(int, const double*, const double*, const double*, const double*, const double*, const double*, const double*,
                            const double*, const double*, int, double*, double*, double*, double*, int*, cudaStream_t);
cudaError_t launch_rts_segment(const double*, const double*, const double*, const double*, const long long*, int, double*, double*, cudaStream_t);
cudaError_t launch_quat_nlerp(const double*, const double*, const double*, long long, double*, cudaStream_t);
cudaError_t launch_sharp_turn(const double*, const double*, const long long*, int, double, int*, double*, cudaStream_t);
struct UtmConst { double A_k0; double e, e2; double alpha[6], beta[6]; double lon0; double fn; };
cudaError_t launch_utm(bool inverse, const double* a, const double* b, long long n, const UtmConst& K, double* o1, double* o2,
                       int num_sms, cudaStream_t stream);
cudaError_t launch_gnss_rows(const double* rows, long long n, const UtmConst& K, double* part, int nparts, double* zone_out,
                             double* out_ts, double* out_xyz, int num_sms, cudaStream_t stream);
cudaError_t launch_geo_mean(const double* lon, const double* lat, long long n, double* part, int nparts, double* out, cudaStream_t stream);
struct AssocArgs {
    const double* gps_t; const double* gps_xyz; const long long* gps_off;
    const double* slam_t; const long long* slam_off;
    double gap; double* aligned; unsigned char* valid; double* work; int B;
};
cudaError_t launch_associate(const AssocArgs& a, int num_sms, cudaStream_t stream);
long long associate_long_work_doubles(long long M, long long N);
cudaError_t launch_associate_long(const double*, const double*, long long, const double*, long long, double, double*, double*, unsigned char*, int*,
                                  cudaStream_t);
int sim3_tiles_for(long long max_len);
cudaError_t launch_umeyama(const double*, const double*, const long long*, const unsigned char*, int, long long, double*,
                           double*, double*, double*, int*, cudaStream_t);
cudaError_t launch_sim3_partial_stats(const double*, const double*, const long long*, const unsigned char*, long long, double*, double*, cudaStream_t);
cudaError_t launch_sim3_from_partial_stats(const double*, int, double*, double*, double*, int*, cudaStream_t);
cudaError_t launch_sim3_ransac(const double*, const double*, long long, const int*, int, int, double, int, double*, unsigned char*,
                               double*, double*, double*, int*, int*, int, cudaStream_t);
long long sim3_ransac_work_doubles(int T, long long n);
cudaError_t launch_sim3_apply(const double*, const double*, const long long*, const double*, const double*, const double*, int,
                              long long, double*, double*, int*, int, cudaStream_t);
struct AteArgs {
    const double* traj; const double* cand; const double* ts; const long long* offsets;
    double skip; double* stats; int B; int cap;
    double* work;             // NULL: candidates + errors in shared memory (cap evaluation poses); else 4 doubles per pose in global memory
};
size_t ate_smem_bytes(int cap);
int ate_smem_capacity(int max_smem);
cudaError_t launch_ate(const AteArgs& a, int num_sms, cudaStream_t stream);
cudaError_t launch_poly_ransac(const double*, const double*, int, const int*, const long long*, const int*, const int*, const int*,
                               const long long*, int, int, int, int, double, unsigned char*, int*, int*, cudaStream_t);
long long parse_table_work_bytes(long long nbytes);
cudaError_t launch_parse_table(const char*, long long, int, int, double*, long long, void*, long long*, cudaStream_t);
long long write_rows_work_bytes(long long n);
cudaError_t launch_write_pose_rows(const double*, const double*, const double*, long long, const int*, const char*, int, char*, long long,
                                   void*, long long*, cudaStream_t);
cudaError_t launch_to_local_f32(const double*, const double*, const double*, const double*, const long long*, const long long*, int, float*, float*,
                                float*, float*, double*, int, cudaStream_t);
cudaError_t launch_from_local_f32(const float*, const float*, const long long*, const long long*, int, const double*, double*, double*, int, cudaStream_t);
cudaError_t launch_fuse_f32(const float*, const float*, const float*, const float*, const double*, const long long*, const long long*, int,
                            const FuseParams*, int, float*, float*, double*, int*, cudaStream_t);
struct SynthArgs {
    double* ts; double* pos; double* quat; double* z;
    long long traj0; int B; int n; double dt; double speed; unsigned long long seed;
    double outage_prob; int outage_max_len;
};
cudaError_t launch_synth(const SynthArgs& a, cudaStream_t stream);
}  // namespace gsf
