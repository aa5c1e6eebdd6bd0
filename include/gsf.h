/* gsf.h -- C ABI of libgsf.so: B200 (sm_100a) kernels for the GPS/SLAM fusion path of
 * A2ureeE/GPS-optimize-SLAM (EKFGPSSLAM.py).
 *
 * The reference is pure Python and has no FFI of its own; the boundary it offers is the
 * function surface of EKFGPSSLAM.py.  Every entry point below names the reference function
 * (file:line in /root/reference/EKFGPSSLAM.py) it replaces; INTEGRATION.md shows the ctypes
 * binding a maintainer would add on the reference side.
 *
 * Conventions
 *   - All arithmetic is IEEE fp64.  Arrays are C-contiguous row-major AoS exactly as the
 *     reference holds them: timestamps [n], positions [n,3], quaternions [n,4] xyzw.
 *   - Batches are ragged: `offsets[B+1]` (int64, in poses) delimits trajectory b as
 *     [offsets[b], offsets[b+1]).
 *   - "_dev" functions take DEVICE pointers and a cudaStream_t (as void*); they launch
 *     asynchronously, allocate nothing and never synchronise.  All device pointers must be
 *     16-byte aligned.  "_host" functions take HOST pointers, stage through an internal
 *     workspace and return when the results are in the host buffers.
 *   - Return value: 0 on success, a negative GSF_E_* code otherwise (gsf_last_error() has
 *     the text).  Per-trajectory conditions are reported in `status[b]` as GSF_ST_* bits.
 *   - A "measurement" row of NaNs means "no GNSS at this pose" (the reference's
 *     valid_mask == False, EKFGPSSLAM.py:867-869).
 */
#ifndef GSF_H
#define GSF_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GSF_E_INVALID   (-1)   /* bad argument (null pointer, misalignment, size) */
#define GSF_E_CUDA      (-2)   /* CUDA runtime error */
#define GSF_E_NO_DEVICE (-3)   /* no usable sm_100 device: there is no CPU fallback */
#define GSF_E_TOO_LARGE (-4)   /* a size beyond what an entry point supports (see its comment) */

#define GSF_ST_OK               0
#define GSF_ST_TOO_FEW_POINTS   1   /* (None,None,None) :431 / ValueError :975,:997 */
#define GSF_ST_DEGENERATE       2   /* collinear points: rotation not unique */
#define GSF_ST_BAD_QUATERNION   4   /* zero-norm quaternion (scipy ValueError, :84, :466) */
#define GSF_ST_EMPTY            8
#define GSF_ST_RANSAC_OUTLIERS 16   /* all-points fit has residuals >= threshold (:411) */
#define GSF_ST_TOO_LONG        32
#define GSF_ST_GRID_NEEDS_ALL_VALID 64 /* hypothesis grid: a pose without GNSS (use gsf_fuse_batched_dev) */
#define GSF_ST_NEEDS_FP64     128   /* fp32 mode: outage / gap / Sim3 window / too few points: use gsf_fuse_batched_dev */

/* Flattened CONFIG (EKFGPSSLAM.py:22-71); 184 bytes, layout fixed. */
typedef struct gsf_fuse_params {
    double p0[7];            /* ekf.initial_cov_diag      (only [0..2] influence outputs) */
    double q[7];             /* ekf.process_noise_diag, per second */
    double r[3];             /* ekf.meas_noise_diag, used as variances (:686) */
    double gap_threshold;    /* time_alignment.max_gps_gap_threshold */
    double max_duration;     /* sim3_ransac.max_initial_duration */
    double yaw_rate_thresh;  /* rts_decision.sharp_turn_yaw_rate_threshold, rad/s */
    double residual_thresh;  /* sim3_ransac.residual_threshold (<= 0: skip the check) */
    double eval_skip;        /* seconds dropped before evaluation (:1021) */
    int32_t min_samples;     /* sim3_ransac.min_samples */
    int32_t sharp_turn_steps;/* rts_decision.default_ekf_transition_steps_on_sharp_turn */
} gsf_fuse_params;

const char* gsf_version(void);
const char* gsf_last_error(void);
/* Number of SMs of the current device, or GSF_E_NO_DEVICE. */
int gsf_device_sm_count(void);

/* ---- fused path: Sim3 point selection (:972-998) + compute_sim3_transform (:428-459) +
 *      row 0 of transform_trajectory (:461-467) + apply_ekf_correction (:831-935), for B
 *      independent trajectories with pre-associated measurements z.
 *      params: 1 record (params_per_traj = 0) or B records.
 *      init_pos/init_quat: NULL, or [B,3]/[B,4] to skip the Sim3 stage and start the filter
 *      from a given pose (the stand-alone apply_ekf_correction contract).
 *      sim3_out [B,16]: R(9, row-major) t(3) s n_selected n_valid n_residual_violators; may be NULL.
 *      Sim3 here is the all-points Umeyama fit.  It equals compute_sim3_transform_robust (:389-426) whenever RANSAC's
 *      best trial keeps every selected point; status bit GSF_ST_RANSAC_OUTLIERS (and n_residual_violators > 0) marks a
 *      trajectory where a residual reaches residual_threshold, i.e. where the reference would have refitted on an
 *      inlier subset: run gsf_sim3_ransac_dev + gsf_sim3_apply_dev and call again with init_pos / init_quat for it.
 *      max_len: largest trajectory length in the batch (sizes the shared-memory staging).  Any length is
 *      accepted: trajectories beyond the shared-memory staging (about 4000 poses) are streamed through
 *      shared memory in tiles by a third kernel (csrc/gsf_long.cu) behind the other two.
 *      Dispatch (invisible to the caller, same results to rounding): batches with max_len <= 1088 and no
 *      init_pos first run the warp-specialised kernel (csrc/gsf_fast.cu); trajectories it does not handle --
 *      a pose without GNSS (outage / RTS :875-928), a GNSS time gap, a time step <= 1e-6 s (the clamp of :863),
 *      a trajectory longer than the Sim3 window, fewer than min_samples points, a zero quaternion, a
 *      measurement-to-process noise ratio outside the range of its projective covariance recursion -- are
 *      processed by the general kernel (csrc/gsf_fused.cu) launched behind it on the same stream.  No host
 *      synchronisation, no allocation; GSF_FUSE_IMPL=general forces the general kernel for everything. */
int gsf_fuse_batched_dev(const double* ts, const double* pos, const double* quat, const double* z,
                         const int64_t* offsets, int32_t B, int64_t max_len,
                         const gsf_fuse_params* params, int32_t params_per_traj,
                         const double* init_pos, const double* init_quat,
                         double* out_pos, double* out_quat, double* sim3_out, int32_t* status,
                         void* stream);

/* ---- noise-parameter hypothesis grid (BASELINE config 5): ONE trajectory, H parameter records; per
 *      hypothesis the EKF of apply_ekf_correction (:831-935) from the Sim3-aligned pose 0 (Sim3 selection
 *      :972-998 and Umeyama :428-459 run once, with params[0]'s gap / window / min_samples), scored with
 *      the evaluation of :1021-1033.  Every pose must carry a GNSS measurement (else status =
 *      GSF_ST_GRID_NEEDS_ALL_VALID and NaN statistics).  stats [H,4]: mean, median, RMSE, count.
 *      sim3_out [13]: R(9) t(3) s, or NULL.  status [1].  work: gsf_hypothesis_grid_work_doubles(). */
int64_t gsf_hypothesis_grid_work_doubles(int64_t n, int32_t H);
int gsf_ekf_hypothesis_grid_dev(const double* ts, const double* pos, const double* quat, const double* z, int64_t n,
                                const gsf_fuse_params* params, int32_t H, double* work, double* stats,
                                double* sim3_out, int32_t* status, void* stream);

/* ---- the same for a PRODUCT grid of noise hypotheses (BASELINE config 5: 64 x 64 x 64 over q_xy, q_z, r): hypothesis
 *      h = (iq * Kz + iz) * Kr + ir runs ExtendedKalmanFilter (:679-772) with process_noise_diag[0:2] = q_xy[iq],
 *      process_noise_diag[2] = q_z[iz], meas_noise_diag = r[ir] (x3) and everything else from `base` (one record).
 *      P0 / Q / R are diagonal (:684-686), so the x / y tracks depend on (q_xy, r) only and the z track on (q_z, r)
 *      only: Kq*Kr + Kq*Kr + Kz*Kr scalar filter runs (the components the per-hypothesis entry above computes, evaluated
 *      parallel in time: equal to rounding) and one nearest-neighbour evaluation (:1021-1033) per hypothesis.  The call scores the hypotheses
 *      [h_first, h_first + h_count) (a rank's shard); stats [h_count,4]: mean, median, RMSE, count.  q_xy [Kq],
 *      q_z [Kz], r [Kr] device fp64.  work: gsf_noise_grid_work_doubles() doubles.  Same restriction and status as
 *      gsf_ekf_hypothesis_grid_dev. */
int64_t gsf_noise_grid_work_doubles(int64_t n, int32_t Kq, int32_t Kz, int32_t Kr, int64_t h_first, int64_t h_count);
int gsf_ekf_noise_grid_dev(const double* ts, const double* pos, const double* quat, const double* z, int64_t n,
                           const gsf_fuse_params* base, const double* q_xy, const double* q_z, const double* r,
                           int32_t Kq, int32_t Kz, int32_t Kr, int64_t h_first, int64_t h_count,
                           double* work, double* stats, double* sim3_out, int32_t* status, void* stream);

/* ---- GNSS outlier pre-filter: the per-window, per-axis polynomial RANSAC of filter_gps_outliers_ransac (:136-247), i.e.
 *      sklearn's RANSACRegressor over PolynomialFeatures(degree) + LinearRegression (absolute loss, `<=` threshold, skip on
 *      fewer inliers, R^2 tie-break, dynamic max_trials), for F independent fits in one launch.  Fit f covers the points
 *      window_idx[fit_offsets[f] .. fit_offsets[f+1]) (indices into t / y) and the column fit_axis[f] of y (row stride
 *      y_stride doubles).  samples [F, max_trials, min_samples]: HOST-SUPPLIED sample indices into the fit's window (what
 *      sklearn draws with sample_without_replacement per trial); dyn_trials[dyn_offsets[f] + k] = sklearn's
 *      _dynamic_max_trials(k, n_window, min_samples, 0.99) for k = 0..n_window, so the loop stops exactly where sklearn's
 *      does.  Outputs: inlier_mask [fit_offsets[F]] (the best trial's), n_trials [F] (draws consumed), status [F]
 *      (1: no consensus set, sklearn's ValueError -> the reference skips the window, :228).  The window enumeration
 *      (:203-238), the AND over the axes and the OR over the windows (:216-219) are index bookkeeping of the caller. */
int gsf_poly_ransac_dev(const double* t, const double* y, int32_t y_stride, const int32_t* window_idx, const int64_t* fit_offsets,
                        const int32_t* fit_axis, const int32_t* samples, const int32_t* dyn_trials, const int64_t* dyn_offsets,
                        int32_t F, int32_t min_samples, int32_t degree, int32_t max_trials, double residual_threshold,
                        uint8_t* inlier_mask, int32_t* n_trials, int32_t* status, void* stream);

/* ---- file formats on the device.
 *      gsf_parse_table_dev: np.loadtxt of a numeric text table already in device memory (load_slam_trajectory :110-125,
 *      load_gps_data :252-258): '#' comments, blank lines skipped, delimiter 0 = runs of whitespace, else exactly that
 *      character (the reference tries ' ' then ',', :252-253).  Every field is converted like strtod (correctly rounded:
 *      the doubles equal numpy's).  out [max_rows, max_cols] row-major, rows in file order.  info [4] (device):
 *      rows, min columns, max columns, status bits (1 unparsable / empty field = numpy's ValueError, 2 a row has more
 *      than max_cols columns, 4 a field has more than 19 significant digits and its rounding is ambiguous, 8 more rows
 *      than max_rows).  Synchronises the stream once (the line count sizes the second half).
 *      gsf_write_pose_rows_dev: the np.savetxt calls of :1087-1102 -- header (host string, written verbatim, may be
 *      NULL), then n rows "ts a b c qx qy qz qw\n" with column k printed as "%.{decimals[k]}f" exactly as printf does
 *      (decimals: host int32 [8]; UTM file 6,6,6,6,8,8,8,8; WGS84 file 6,8,8,3,8,8,8,8).  out_info [2] (device): bytes
 *      written, 1 if a value was outside the formatter's range (|x| >= 2^63). */
int64_t gsf_parse_table_work_bytes(int64_t nbytes);
int gsf_parse_table_dev(const char* text, int64_t nbytes, int32_t delimiter, int32_t max_cols, double* out, int64_t max_rows,
                        void* work, int64_t* info, void* stream);
int64_t gsf_write_rows_work_bytes(int64_t n);
int gsf_write_pose_rows_dev(const double* ts, const double* xyz, const double* quat, int64_t n, const int32_t* decimals,
                            const char* header, int32_t header_bytes, char* out, int64_t capacity, void* work, int64_t* out_info,
                            void* stream);

/* ---- optional fp32 mode of the fused path (same stages as gsf_fuse_batched_dev; target 1e-4 m).  Storage is fp32 relative
 *      to per-trajectory fp64 origins -- origins [B,7] = t0, SLAM origin (3), UTM origin (3); ts32 = ts - t0,
 *      pos32 = pos - SLAM origin, z32 = z - UTM origin (NaN = no GNSS), quat32; outputs out_pos32 (fused position - UTM
 *      origin), out_quat32 -- and INTERLEAVED BY 32 TRAJECTORIES: trajectories 32 g .. 32 g + 31 form group g, padded to
 *      its longest member; group_offsets [ceil(B/32) + 1] (int64) = running sum of the padded lengths; element (pose i,
 *      component c, member l) of a W-component array lives at ((group_offsets[g] + i) * W + c) * 32 + l.  A warp is a
 *      group, so every access is one 128-byte line.  44 B in + 28 B out per pose instead of 88 + 56.
 *      Reductions (Umeyama sums), the SVD and pose 0 are fp64; the filter runs in fp32 in innovation form (state minus
 *      measurement), so its rounding is 1e-7 of metres and the result is limited by the fp32 storage of the relative
 *      coordinates (6e-5 m at 1 km from the origin).  sim3_out [B,16] as in gsf_fuse_batched_dev, t in the absolute
 *      frames.  One thread per trajectory.  Trajectories that need the general machinery (a pose without GNSS, a GNSS gap,
 *      the Sim3 window, a step <= 1e-6 s, too few points, a zero first quaternion) return GSF_ST_NEEDS_FP64 and NaN rows.
 *      gsf_to_local_f32_dev / gsf_from_local_f32_dev convert fp64 AoS arrays (offsets [B+1]) to the format and the fused
 *      positions / quaternions back (origin: the middle pose; `pos` or `quat` may be NULL in the second). */
int gsf_to_local_f32_dev(const double* ts, const double* pos, const double* quat, const double* z, const int64_t* offsets,
                         const int64_t* group_offsets, int32_t B, float* ts32, float* pos32, float* quat32, float* z32, double* origins,
                         void* stream);
int gsf_from_local_f32_dev(const float* pos32, const float* quat32, const int64_t* offsets, const int64_t* group_offsets, int32_t B,
                           const double* origins, double* pos, double* quat, void* stream);
int gsf_fuse_batched_f32_dev(const float* ts32, const float* pos32, const float* quat32, const float* z32, const double* origins,
                             const int64_t* offsets, const int64_t* group_offsets, int32_t B, const gsf_fuse_params* params,
                             int32_t params_per_traj, float* out_pos32, float* out_quat32, double* sim3_out, int32_t* status, void* stream);

/* ---- apply_ekf_correction (:831-935), literal step-by-step recursion, one thread per
 *      trajectory (general path; keeps the zero-motion fallback of :84-86). */
int gsf_ekf_strict_batched_dev(const double* ts, const double* pos, const double* quat, const double* z,
                               const int64_t* offsets, int32_t B,
                               const gsf_fuse_params* params, int32_t params_per_traj,
                               const double* init_pos, const double* init_quat,
                               double* out_pos, double* out_quat, int32_t* status, void* stream);

/* ---- per-call surface of the reference's EKF helpers, dense 7x7 covariance as the Python objects hold it, batched
 *      over B independent filters / segments (API completeness; the throughput path is gsf_fuse_batched_dev).
 *      gsf_ekf_step_dev: ExtendedKalmanFilter._predict (:702-715) [mode bit 0], ._update (:717-734) [mode bit 1] and
 *      the blend of process_step (:754-767) [blend_w < 1].  state [B,7], cov [B,49], motion_dp [B,3], motion_dq [B,4],
 *      dt [B], z [B,3] (NaN row: no update), q_diag_per_sec [7], r_diag [3], blend_w [B].  Outputs: out_state, out_cov,
 *      pred_state, pred_cov, flags [B] (1 update applied, 2 zero-norm quaternion = scipy ValueError, 4 singular S).
 *      gsf_rts_segment_dev: rts_smoother_segment (:777-803) over segments [offsets[b], offsets[b+1]).
 *      gsf_quat_nlerp_dev: quaternion_nlerp (:94-105).  gsf_sharp_turn_dev: is_sharp_turn_in_segment (:808-826),
 *      flags [B] and the largest yaw rate [B] (rad/s). */
int gsf_ekf_step_dev(int32_t mode, const double* state, const double* cov, const double* motion_dp, const double* motion_dq,
                     const double* dt, const double* z, const double* q_diag_per_sec, const double* r_diag,
                     const double* blend_w, int32_t B, double* out_state, double* out_cov, double* pred_state,
                     double* pred_cov, int32_t* flags, void* stream);
int gsf_rts_segment_dev(const double* x_filt, const double* P_filt, const double* x_pred, const double* P_pred,
                        const int64_t* offsets, int32_t B, double* x_smooth, double* P_smooth, void* stream);
int gsf_quat_nlerp_dev(const double* q1, const double* q2, const double* weight_q2, int64_t n, double* out, void* stream);
int gsf_sharp_turn_dev(const double* ts, const double* quat, const int64_t* offsets, int32_t B, double yaw_rate_threshold,
                       int32_t* flags, double* max_rate, void* stream);

/* ---- compute_sim3_transform (:428-459) batched.  mask: NULL or one byte per point.
 *      work: gsf_umeyama_work_doubles(B, max_len) doubles.  R [B,9], t [B,3], s [B]. */
int64_t gsf_umeyama_work_doubles(int32_t B, int64_t max_len);
int gsf_sim3_umeyama_batched_dev(const double* src, const double* dst, const int64_t* offsets,
                                 const uint8_t* mask, int32_t B, int64_t max_len, double* work,
                                 double* R, double* t, double* s, int32_t* status, void* stream);

/* ---- compute_sim3_transform (:428-459) for ONE trajectory whose points are spread over several GPUs (BASELINE config 4
 *      sharded): every rank reduces its shard to GSF_SIM3_STATS doubles -- count, mean src [3], mean dst [3],
 *      centred cross-covariance [9] (not / n), sum |src - mean|^2, 3 of padding -- the ranks exchange them (an all-gather
 *      of 160 bytes each) and every rank merges them in shard order (pairwise covariance update, deterministic) and finishes.
 *      offsets2 = {0, n} on the device; mask NULL or one byte per point; work: gsf_umeyama_work_doubles(1, n) doubles.
 *      stats [shards, GSF_SIM3_STATS]; R [9], t [3], s [1], status [1]. */
#define GSF_SIM3_STATS 20
int gsf_sim3_partial_stats_dev(const double* src, const double* dst, const int64_t* offsets2, const uint8_t* mask, int64_t n,
                               double* work, double* stats, void* stream);
int gsf_sim3_from_partial_stats_dev(const double* stats, int32_t shards, double* R, double* t, double* s, int32_t* status, void* stream);

/* ---- compute_sim3_transform_robust (:389-426) with HOST-SUPPLIED sample indices: `samples`
 *      [trials, min_samples] int32 (device) holds what the reference draws with
 *      np.random.choice(n, min_samples, replace=False) in each trial (:408), so a seeded reference
 *      run and this call evaluate the same trials.  Per trial: Umeyama on the sample, residual
 *      norms of all n points, inlier count; first strictly-best trial wins (:416); final fit on
 *      its inliers (:422-423).  inlier_mask [n]; info[3] = max_inliers, winning trial, usable flag
 *      (max_inliers >= min_inliers, :419); R [9], t [3], s [1]; status[1] = GSF_ST_TOO_FEW_POINTS
 *      when the reference returns (None, None, None).  work: gsf_sim3_ransac_work_doubles(). */
int64_t gsf_sim3_ransac_work_doubles(int32_t trials, int64_t n);
int gsf_sim3_ransac_dev(const double* src, const double* dst, int64_t n, const int32_t* samples,
                        int32_t trials, int32_t min_samples, double residual_threshold, int32_t min_inliers,
                        double* work, uint8_t* inlier_mask, double* R, double* t, double* s,
                        int32_t* info, int32_t* status, void* stream);

/* ---- transform_trajectory (:461-467) batched; status must be zero-initialised. */
int gsf_sim3_apply_dev(const double* pos, const double* quat, const int64_t* offsets,
                       const double* R, const double* t, const double* s, int32_t B, int64_t max_len,
                       double* out_pos, double* out_quat, int32_t* status, void* stream);

/* ---- evaluation (:1021-1033): nearest-neighbour error statistics of `traj` against the
 *      candidate set {cand[i] : cand[i] not NaN and ts[i] > ts[0] + skip}, exact (bucketed, pruned search
 *      instead of the reference's full cdist matrix).  stats [B,4]: mean, median, RMSE, count.
 *      Any trajectory length: up to ~6900 poses the candidates and errors of a trajectory are held in
 *      shared memory and `work` may be NULL; beyond that `work` must hold gsf_ate_work_doubles(total
 *      poses of the batch, max_len) doubles (0 when shared memory suffices). */
int64_t gsf_ate_work_doubles(int64_t total_poses, int64_t max_len);
int gsf_ate_nn_batched_dev(const double* traj, const double* cand, const double* ts,
                           const int64_t* offsets, int32_t B, int64_t max_len, double skip,
                           double* work, double* stats, void* stream);

/* ---- pyproj Proj("+proj=utm +zone=Z [+south] +ellps=WGS84") forward (:270) / inverse (:295). */
int gsf_utm_forward_dev(const double* lon, const double* lat, int64_t n, int32_t zone, int32_t south,
                        double* east, double* north, void* stream);
int gsf_utm_inverse_dev(const double* east, const double* north, int64_t n, int32_t zone, int32_t south,
                        double* lon, double* lat, void* stream);
/* ---- auto_utm_projection (:127-134) on the device: out[4] = mean lon, mean lat, zone, south.
 *      part: 2*GSF_GEO_PARTS doubles of scratch. */
#define GSF_GEO_PARTS 1024
int gsf_geo_zone_dev(const double* lon, const double* lat, int64_t n, double* part, double* out, void* stream);

/* ---- fused GNSS ingest: the projection part of load_gps_data (:258-271) with auto_utm_projection
 *      (:127-134).  rows [n,4] = timestamp, lat, lon, alt in the loader's column order (:258), 16-byte
 *      aligned.  Rows failing the validity mask (:259) do not enter the zone means and come out as NaN
 *      measurements.  zone_out[5] = mean lon, mean lat, zone, south flag, valid rows (device; the
 *      projection reads the zone from there: no host round trip; it runs once with the zone of the first
 *      4096 rows while the means accumulate, and again only if the whole-track zone differs).
 *      part: 3*GSF_GEO_PARTS doubles.
 *      out_ts [n] (may be NULL), out_xyz [n,3] = easting, northing, altitude. */
int gsf_gnss_rows_to_utm_dev(const double* rows, int64_t n, double* part, double* zone_out,
                             double* out_ts, double* out_xyz, void* stream);

/* ---- dynamic_time_alignment (:325-387) after the host-side argsort/unique (:340-349):
 *      per-segment not-a-knot cubic / linear interpolation of the GNSS track at the SLAM
 *      stamps.  work: 4 doubles per GNSS sample.  aligned [n_slam,3] NaN-filled, valid [n_slam]. */
int gsf_associate_spline_dev(const double* gps_t, const double* gps_xyz, const int64_t* gps_offsets,
                             const double* slam_t, const int64_t* slam_offsets, int32_t B, double gap,
                             double* work, double* aligned, uint8_t* valid, void* stream);

/* ---- the same for ONE trajectory of any size (BASELINE config 4: 1e8 samples): the not-a-knot system is solved locally --
 *      13-knot chunks with a 20-knot halo on both sides, natural ends where the halo cuts a segment, the true end rows where
 *      the segment ends inside it; the cut decays like 0.268^20 = 4e-12 of a centimetre-sized curvature term, below fp64
 *      rounding of the coordinates -- and the evaluation searches each SLAM stamp inside the knot bracket of its tile of
 *      2048 stamps.  gps_t [M] sorted and unique, gps_xyz [M,3], slam_t [N] (any order; sorted stamps keep the brackets
 *      small).  work: gsf_associate_spline_long_work_doubles(M, N) doubles (3 M + 2 + 2 ceil(N / 2048)).
 *      status [1] (may be NULL): 1 if two consecutive GNSS stamps differ by <= 1e-9 s inside a segment (the reference
 *      drops such a segment, :356-359; use gsf_associate_spline_dev for that data). */
int64_t gsf_associate_spline_long_work_doubles(int64_t M, int64_t N);
int gsf_associate_spline_long_dev(const double* gps_t, const double* gps_xyz, int64_t M, const double* slam_t, int64_t N, double gap,
                                  double* work, double* aligned, uint8_t* valid, int32_t* status, void* stream);

/* ---- synthetic batch generator for the large bench configs (equal lengths n). */
int gsf_synth_generate_dev(double* ts, double* pos, double* quat, double* z, int64_t first_traj,
                           int32_t B, int32_t n, double dt, double speed, uint64_t seed,
                           double outage_prob, int32_t outage_max_len, void* stream);

/* ---- host-buffer entry point of the fused path (the call a reference-side binding makes):
 *      same contract as gsf_fuse_batched_dev with HOST pointers; copies in, runs, copies out,
 *      chunked and double-buffered on internal streams.  Returns when outputs are complete. */
int gsf_fuse_batched_host(const double* ts, const double* pos, const double* quat, const double* z,
                          const int64_t* offsets, int32_t B, int64_t max_len,
                          const gsf_fuse_params* params, int32_t params_per_traj,
                          const double* init_pos, const double* init_quat,
                          double* out_pos, double* out_quat, double* sim3_out, int32_t* status);
/* Release the internal workspace of the host entry points. */
void gsf_host_workspace_free(void);

#ifdef __cplusplus
}
#endif
#endif /* GSF_H */
