"""Batched device API of the fusion path: thin wrappers that hand torch CUDA tensors
(device memory + streams only) to the C ABI of libgsf.so.  Every function raises if the
library or a B200 is missing -- nothing here computes on the CPU.

Layout: fp64, C-contiguous, ragged batches described by ``offsets`` (int64 [B+1], in poses):
ts [P], pos [P,3], quat [P,4] (xyzw), z [P,3] (NaN row = no GNSS at that pose).
"""
from __future__ import annotations

import ctypes

import numpy as np
import torch

from . import _lib
from .config import CONFIG, pack_fuse_params


def _ptr(t):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def _stream_ptr(stream=None):
    s = torch.cuda.current_stream() if stream is None else stream
    return ctypes.c_void_p(s.cuda_stream)


def _f64(t, device=None):
    if isinstance(t, np.ndarray):
        t = torch.from_numpy(np.ascontiguousarray(t, dtype=np.float64))
    t = t.to(device=device or "cuda", dtype=torch.float64)
    return t.contiguous()


def _require_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise _lib.GsfError("device API called with a CPU tensor (no CPU fallback exists)")


def equal_offsets(batch: int, n: int, device="cuda") -> torch.Tensor:
    return torch.arange(batch + 1, dtype=torch.int64, device=device) * n


def params_tensor(cfg=None, device="cuda", per_traj=None) -> torch.Tensor:
    """One packed FuseParams record (uint8[184]) or ``per_traj`` = list of (p0, q, r) overrides."""
    if per_traj is None:
        blob = pack_fuse_params(cfg)
    else:
        blob = np.concatenate([pack_fuse_params(cfg, p0=p0, q=q, r=r) for (p0, q, r) in per_traj])
    return torch.from_numpy(blob).to(device)


def fuse_batched(ts, pos, quat, z, offsets, max_len, params, params_per_traj=False,
                 init_pos=None, init_quat=None, out_pos=None, out_quat=None, sim3_out=None,
                 status=None, stream=None):
    """Sim3 selection + Umeyama + EKF/RTS for B trajectories (gsf_fuse_batched_dev).
    Returns (out_pos [P,3], out_quat [P,4], sim3 [B,16], status [B] int32); asynchronous."""
    lib = _lib.load()
    _require_cuda(ts, pos, quat, z, offsets, params)
    B = offsets.numel() - 1
    P = ts.numel()
    dev = ts.device
    if out_pos is None:
        out_pos = torch.empty((P, 3), dtype=torch.float64, device=dev)
    if out_quat is None:
        out_quat = torch.empty((P, 4), dtype=torch.float64, device=dev)
    if sim3_out is None:
        sim3_out = torch.empty((B, 16), dtype=torch.float64, device=dev)
    if status is None:
        status = torch.empty((B,), dtype=torch.int32, device=dev)
    rc = lib.gsf_fuse_batched_dev(_ptr(ts), _ptr(pos), _ptr(quat), _ptr(z), _ptr(offsets), B, int(max_len),
                                  _ptr(params), int(bool(params_per_traj)), _ptr(init_pos), _ptr(init_quat),
                                  _ptr(out_pos), _ptr(out_quat), _ptr(sim3_out), _ptr(status), _stream_ptr(stream))
    _lib.check(rc, "gsf_fuse_batched_dev")
    return out_pos, out_quat, sim3_out, status


def ekf_step(mode, state, cov, motion_dp, motion_dq, dt, z, q_diag, r_diag, blend_w, stream=None):
    """ExtendedKalmanFilter._predict / ._update / blend for B filters (gsf_ekf_step_dev); device fp64 tensors.
    Returns (out_state [B,7], out_cov [B,7,7], pred_state [B,7], pred_cov [B,7,7], flags [B] int32)."""
    lib = _lib.load()
    _require_cuda(state, cov, motion_dp, motion_dq, dt, z, q_diag, r_diag, blend_w)
    B = state.shape[0]
    dev = state.device
    out_state = torch.empty((B, 7), dtype=torch.float64, device=dev); out_cov = torch.empty((B, 7, 7), dtype=torch.float64, device=dev)
    pred_state = torch.empty((B, 7), dtype=torch.float64, device=dev); pred_cov = torch.empty((B, 7, 7), dtype=torch.float64, device=dev)
    flags = torch.empty((B,), dtype=torch.int32, device=dev)
    rc = lib.gsf_ekf_step_dev(int(mode), _ptr(state), _ptr(cov), _ptr(motion_dp), _ptr(motion_dq), _ptr(dt), _ptr(z), _ptr(q_diag),
                              _ptr(r_diag), _ptr(blend_w), int(B), _ptr(out_state), _ptr(out_cov), _ptr(pred_state), _ptr(pred_cov),
                              _ptr(flags), _stream_ptr(stream))
    _lib.check(rc, "gsf_ekf_step_dev")
    return out_state, out_cov, pred_state, pred_cov, flags


def rts_segments(x_filt, P_filt, x_pred, P_pred, offsets, stream=None):
    """rts_smoother_segment for B segments (gsf_rts_segment_dev) -> (x_smooth [N,7], P_smooth [N,7,7])."""
    lib = _lib.load()
    _require_cuda(x_filt, P_filt, x_pred, P_pred, offsets)
    xs = torch.empty_like(x_filt); Ps = torch.empty_like(P_filt)
    rc = lib.gsf_rts_segment_dev(_ptr(x_filt), _ptr(P_filt), _ptr(x_pred), _ptr(P_pred), _ptr(offsets), int(offsets.numel() - 1),
                                 _ptr(xs), _ptr(Ps), _stream_ptr(stream))
    _lib.check(rc, "gsf_rts_segment_dev")
    return xs, Ps


def quat_nlerp(q1, q2, w, stream=None):
    """quaternion_nlerp for n pairs (gsf_quat_nlerp_dev)."""
    lib = _lib.load()
    _require_cuda(q1, q2, w)
    out = torch.empty_like(q1)
    rc = lib.gsf_quat_nlerp_dev(_ptr(q1), _ptr(q2), _ptr(w), int(q1.shape[0]), _ptr(out), _stream_ptr(stream))
    _lib.check(rc, "gsf_quat_nlerp_dev")
    return out


def sharp_turn(ts, quat, offsets, thresh, stream=None):
    """is_sharp_turn_in_segment for B segments (gsf_sharp_turn_dev) -> (flags [B] int32, max yaw rate [B])."""
    lib = _lib.load()
    _require_cuda(ts, quat, offsets)
    B = offsets.numel() - 1
    flags = torch.empty((B,), dtype=torch.int32, device=ts.device); rate = torch.empty((B,), dtype=torch.float64, device=ts.device)
    rc = lib.gsf_sharp_turn_dev(_ptr(ts), _ptr(quat), _ptr(offsets), int(B), float(thresh), _ptr(flags), _ptr(rate), _stream_ptr(stream))
    _lib.check(rc, "gsf_sharp_turn_dev")
    return flags, rate


def gnss_rows_to_utm(rows, want_ts=True, stream=None, out_ts=None, out_xyz=None):
    """Fused GNSS ingest (gsf_gnss_rows_to_utm_dev): rows [n,4] = ts, lat, lon, alt ->
    (ts [n] or None, xyz [n,3] = E, N, alt with NaN rows where the validity mask fails,
    zone [5] = mean lon, mean lat, zone, south, valid count); asynchronous, zone stays on the device."""
    lib = _lib.load()
    _require_cuda(rows)
    n = rows.shape[0]
    dev = rows.device
    part = torch.empty((3 * _lib.GEO_PARTS,), dtype=torch.float64, device=dev)
    zone = torch.empty((5,), dtype=torch.float64, device=dev)
    ts = out_ts if out_ts is not None else (torch.empty((n,), dtype=torch.float64, device=dev) if want_ts else None)
    xyz = out_xyz if out_xyz is not None else torch.empty((n, 3), dtype=torch.float64, device=dev)      # (contiguous views of a caller's buffer)
    rc = lib.gsf_gnss_rows_to_utm_dev(_ptr(rows), int(n), _ptr(part), _ptr(zone), _ptr(ts), _ptr(xyz), _stream_ptr(stream))
    _lib.check(rc, "gsf_gnss_rows_to_utm_dev")
    return ts, xyz, zone


def hypothesis_grid(ts, pos, quat, z, params, stream=None):
    """One trajectory x H noise-parameter hypotheses -> ATE statistics (gsf_ekf_hypothesis_grid_dev).
    ts [n], pos [n,3], quat [n,4], z [n,3] (all valid); params: uint8 [H*184] packed FuseParams records.
    Returns (stats [H,4] = mean, median, RMSE, count; sim3 [13]; status [1]); asynchronous."""
    lib = _lib.load()
    _require_cuda(ts, pos, quat, z, params)
    n = ts.numel()
    H = params.numel() // 184
    dev = ts.device
    work = torch.empty((lib.gsf_hypothesis_grid_work_doubles(int(n), int(H)),), dtype=torch.float64, device=dev)
    stats = torch.empty((H, 4), dtype=torch.float64, device=dev)
    sim3 = torch.empty((13,), dtype=torch.float64, device=dev)
    status = torch.empty((1,), dtype=torch.int32, device=dev)
    rc = lib.gsf_ekf_hypothesis_grid_dev(_ptr(ts), _ptr(pos), _ptr(quat), _ptr(z), int(n), _ptr(params), int(H), _ptr(work),
                                         _ptr(stats), _ptr(sim3), _ptr(status), _stream_ptr(stream))
    _lib.check(rc, "gsf_ekf_hypothesis_grid_dev")
    return stats, sim3, status


def noise_grid(ts, pos, quat, z, base_params, q_xy, q_z, r, h_first=0, h_count=None, work=None, stats=None, stream=None):
    """One trajectory x a product grid of noise hypotheses (gsf_ekf_noise_grid_dev): hypothesis
    h = (iq * Kz + iz) * Kr + ir uses process noise (q_xy[iq], q_xy[iq], q_z[iz]) and measurement noise r[ir]; everything
    else comes from ``base_params`` (uint8 [184]).  Scores [h_first, h_first + h_count).
    Returns (stats [h_count,4] = mean, median, RMSE, count; sim3 [13]; status [1]); asynchronous."""
    lib = _lib.load()
    _require_cuda(ts, pos, quat, z, base_params, q_xy, q_z, r)
    n = int(ts.numel())
    Kq, Kz, Kr = int(q_xy.numel()), int(q_z.numel()), int(r.numel())
    if h_count is None:
        h_count = Kq * Kz * Kr - h_first
    dev = ts.device
    need = lib.gsf_noise_grid_work_doubles(n, Kq, Kz, Kr, int(h_first), int(h_count))
    if need <= 0:
        raise ValueError("noise_grid: empty or out-of-range hypothesis range")
    if work is None or work.numel() < need:
        work = torch.empty((need,), dtype=torch.float64, device=dev)
    if stats is None:
        stats = torch.empty((h_count, 4), dtype=torch.float64, device=dev)
    sim3 = torch.empty((13,), dtype=torch.float64, device=dev)
    status = torch.empty((1,), dtype=torch.int32, device=dev)
    rc = lib.gsf_ekf_noise_grid_dev(_ptr(ts), _ptr(pos), _ptr(quat), _ptr(z), n, _ptr(base_params), _ptr(q_xy), _ptr(q_z), _ptr(r),
                                    Kq, Kz, Kr, int(h_first), int(h_count), _ptr(work), _ptr(stats), _ptr(sim3), _ptr(status),
                                    _stream_ptr(stream))
    _lib.check(rc, "gsf_ekf_noise_grid_dev")
    return stats, sim3, status


def poly_ransac(t, y, window_idx, fit_offsets, fit_axis, samples, dyn_trials, dyn_offsets, min_samples, degree, max_trials,
                residual_threshold, stream=None):
    """Per-window polynomial RANSAC of the GNSS pre-filter (gsf_poly_ransac_dev, EKFGPSSLAM.py:136-247) for F fits.
    t [N], y [N,C] device fp64; window_idx int32, fit_offsets int64 [F+1], fit_axis int32 [F], samples int32
    [F, max_trials, min_samples], dyn_trials int32, dyn_offsets int64 [F] (all device).
    Returns (inlier_mask uint8 [len(window_idx)], n_trials int32 [F], status int32 [F]); asynchronous."""
    lib = _lib.load()
    _require_cuda(t, y, window_idx, fit_offsets, fit_axis, samples, dyn_trials, dyn_offsets)
    F = int(fit_axis.numel())
    dev = t.device
    mask = torch.empty((int(window_idx.numel()),), dtype=torch.uint8, device=dev)
    n_trials = torch.empty((F,), dtype=torch.int32, device=dev)
    status = torch.empty((F,), dtype=torch.int32, device=dev)
    rc = lib.gsf_poly_ransac_dev(_ptr(t), _ptr(y), int(y.shape[1]) if y.dim() > 1 else 1, _ptr(window_idx), _ptr(fit_offsets), _ptr(fit_axis),
                                 _ptr(samples), _ptr(dyn_trials), _ptr(dyn_offsets), F, int(min_samples), int(degree), int(max_trials),
                                 float(residual_threshold), _ptr(mask), _ptr(n_trials), _ptr(status), _stream_ptr(stream))
    _lib.check(rc, "gsf_poly_ransac_dev")
    return mask, n_trials, status


def dynamic_max_trials_table(n_window: int, min_samples: int, max_trials: int, probability: float = 0.99):
    """sklearn's _dynamic_max_trials(n_inliers, n_window, min_samples, probability) for n_inliers = 0..n_window
    (linear_model/_ransac.py), clipped to max_trials; int32 numpy array."""
    import numpy as np
    eps = np.spacing(1)
    out = np.empty(n_window + 1, dtype=np.int32)
    for k in range(n_window + 1):
        ratio = k / float(n_window)
        nom = max(eps, 1 - probability)
        denom = max(eps, 1 - ratio ** min_samples)
        if nom == 1:
            v = 0.0
        elif denom == 1:
            v = float("inf")
        else:
            v = abs(float(np.ceil(np.log(nom) / np.log(denom))))
        out[k] = int(min(v, max_trials))
    return out


def parse_table(text, delimiter: int = 0, max_cols: int = 16, max_rows=None, stream=None):
    """np.loadtxt of a numeric text table on the device (gsf_parse_table_dev).  text: uint8 CUDA tensor with the file's
    bytes; delimiter 0 = whitespace runs, else ord(char).  Returns (table [rows, cols] float64 CUDA, status bits)
    -- synchronous (the row count comes back to the host)."""
    lib = _lib.load()
    _require_cuda(text)
    nbytes = int(text.numel())
    dev = text.device
    if max_rows is None:
        max_rows = nbytes // 2 + 1                   # a row needs a digit and a newline
    out = torch.empty((max_rows, max_cols), dtype=torch.float64, device=dev)
    work = torch.empty((lib.gsf_parse_table_work_bytes(nbytes),), dtype=torch.uint8, device=dev)
    info = torch.zeros((4,), dtype=torch.int64, device=dev)
    rc = lib.gsf_parse_table_dev(_ptr(text), nbytes, int(delimiter), int(max_cols), _ptr(out), int(max_rows), _ptr(work), _ptr(info),
                                 _stream_ptr(stream))
    _lib.check(rc, "gsf_parse_table_dev")
    rows, cmin, cmax, status = [int(v) for v in info.cpu()]
    if rows > 0 and cmin != cmax:
        status |= 16                                 # ragged rows: numpy's "Wrong number of columns"
    cols = min(cmax, max_cols) if rows > 0 else 0
    return out[:min(rows, max_rows), :cols], status


def write_pose_rows(ts, xyz, quat, decimals, header: str = "", stream=None):
    """np.savetxt of rows ``ts a b c qx qy qz qw`` with "%.Df" per column (gsf_write_pose_rows_dev, EKFGPSSLAM.py:1087-1102).
    Returns the file's bytes as a uint8 CUDA tensor -- synchronous (the length comes back to the host)."""
    import numpy as np
    lib = _lib.load()
    _require_cuda(ts, xyz, quat)
    n = int(ts.numel())
    dev = ts.device
    hdr = header.encode()
    dec = np.ascontiguousarray(decimals, dtype=np.int32)
    if dec.shape != (8,):
        raise ValueError("decimals must have 8 entries")
    cap = len(hdr) + n * (8 * 31) + 16
    out = torch.empty((cap,), dtype=torch.uint8, device=dev)
    work = torch.empty((lib.gsf_write_rows_work_bytes(n),), dtype=torch.uint8, device=dev)
    info = torch.zeros((2,), dtype=torch.int64, device=dev)
    rc = lib.gsf_write_pose_rows_dev(_ptr(ts), _ptr(xyz), _ptr(quat), n, ctypes.c_void_p(dec.ctypes.data), hdr, len(hdr), _ptr(out), cap,
                                     _ptr(work), _ptr(info), _stream_ptr(stream))
    _lib.check(rc, "gsf_write_pose_rows_dev")
    nbytes, bad = [int(v) for v in info.cpu()]
    if bad:
        raise _lib.GsfError("write_pose_rows: a value is outside the fixed-point formatter's range (|x| >= 2^63)")
    return out[:nbytes]


class F32Batch:
    """The fp32 mode's storage of a batch (see include/gsf.h: fp32 relative to fp64 origins, interleaved by 32 trajectories)."""

    def __init__(self, ts32, pos32, quat32, z32, origins, offsets, group_offsets):
        self.ts32, self.pos32, self.quat32, self.z32 = ts32, pos32, quat32, z32
        self.origins, self.offsets, self.group_offsets = origins, offsets, group_offsets

    @property
    def B(self):
        return self.offsets.numel() - 1


def f32_group_offsets(offsets):
    """offsets [B+1] (device int64) -> group_offsets [ceil(B/32)+1]: running sum of every 32-trajectory group's longest length."""
    lengths = offsets[1:] - offsets[:-1]
    B = lengths.numel()
    G = (B + 31) // 32
    padded = torch.zeros((G * 32,), dtype=torch.int64, device=offsets.device)
    padded[:B] = lengths
    Lg = padded.reshape(G, 32).max(dim=1).values
    return torch.cat([torch.zeros((1,), dtype=torch.int64, device=offsets.device), torch.cumsum(Lg, 0)])


def to_local_f32(ts, pos, quat, z, offsets, stream=None):
    """fp64 AoS arrays -> the fp32 mode's storage (gsf_to_local_f32_dev).  Returns an F32Batch (one host sync: the padded size)."""
    lib = _lib.load()
    _require_cuda(ts, pos, quat, z, offsets)
    B = offsets.numel() - 1
    dev = ts.device
    go = f32_group_offsets(offsets)
    rows = int(go[-1].item())
    ts32 = torch.empty((rows * 32,), dtype=torch.float32, device=dev); pos32 = torch.empty((rows * 96,), dtype=torch.float32, device=dev)
    quat32 = torch.empty((rows * 128,), dtype=torch.float32, device=dev); z32 = torch.empty((rows * 96,), dtype=torch.float32, device=dev)
    origins = torch.zeros((B, 7), dtype=torch.float64, device=dev)
    rc = lib.gsf_to_local_f32_dev(_ptr(ts), _ptr(pos), _ptr(quat), _ptr(z), _ptr(offsets), _ptr(go), B, _ptr(ts32), _ptr(pos32), _ptr(quat32),
                                  _ptr(z32), _ptr(origins), _stream_ptr(stream))
    _lib.check(rc, "gsf_to_local_f32_dev")
    return F32Batch(ts32, pos32, quat32, z32, origins, offsets, go)


def from_local_f32(batch, pos32, quat32=None, stream=None):
    """fused positions (and quaternions) of the fp32 mode back to fp64 AoS arrays (gsf_from_local_f32_dev)."""
    lib = _lib.load()
    _require_cuda(pos32)
    P = int(batch.offsets[-1].item())
    dev = pos32.device
    out_p = torch.empty((P, 3), dtype=torch.float64, device=dev)
    out_q = torch.empty((P, 4), dtype=torch.float64, device=dev) if quat32 is not None else None
    rc = lib.gsf_from_local_f32_dev(_ptr(pos32), _ptr(quat32), _ptr(batch.offsets), _ptr(batch.group_offsets), batch.B, _ptr(batch.origins),
                                    _ptr(out_p), _ptr(out_q), _stream_ptr(stream))
    _lib.check(rc, "gsf_from_local_f32_dev")
    return out_p if quat32 is None else (out_p, out_q)


def fuse_batched_f32(batch, params, params_per_traj=False, out_pos=None, out_quat=None, sim3_out=None, status=None, stream=None):
    """Optional fp32 mode of the fused path (gsf_fuse_batched_f32_dev) on an F32Batch.
    Returns (out_pos32, out_quat32 in the batch's interleaved layout, sim3 [B,16], status [B]); asynchronous."""
    lib = _lib.load()
    _require_cuda(batch.ts32, params)
    B = batch.B
    dev = batch.ts32.device
    out_pos = torch.empty_like(batch.pos32) if out_pos is None else out_pos
    out_quat = torch.empty_like(batch.quat32) if out_quat is None else out_quat
    sim3_out = torch.empty((B, 16), dtype=torch.float64, device=dev) if sim3_out is None else sim3_out
    status = torch.empty((B,), dtype=torch.int32, device=dev) if status is None else status
    rc = lib.gsf_fuse_batched_f32_dev(_ptr(batch.ts32), _ptr(batch.pos32), _ptr(batch.quat32), _ptr(batch.z32), _ptr(batch.origins),
                                      _ptr(batch.offsets), _ptr(batch.group_offsets), B, _ptr(params), int(bool(params_per_traj)),
                                      _ptr(out_pos), _ptr(out_quat), _ptr(sim3_out), _ptr(status), _stream_ptr(stream))
    _lib.check(rc, "gsf_fuse_batched_f32_dev")
    return out_pos, out_quat, sim3_out, status


def associate_spline_long(gps_t, gps_xyz, slam_t, gap, stream=None):
    """dynamic_time_alignment for one trajectory of any size (gsf_associate_spline_long_dev): local-halo spline solve.
    Returns (aligned [N,3], valid [N] uint8, status [1] int32); asynchronous."""
    lib = _lib.load()
    _require_cuda(gps_t, gps_xyz, slam_t)
    M, N = int(gps_t.numel()), int(slam_t.numel())
    dev = gps_t.device
    work = torch.empty((int(lib.gsf_associate_spline_long_work_doubles(M, N)),), dtype=torch.float64, device=dev)
    aligned = torch.empty((N, 3), dtype=torch.float64, device=dev)
    valid = torch.empty((N,), dtype=torch.uint8, device=dev)
    status = torch.zeros((1,), dtype=torch.int32, device=dev)
    rc = lib.gsf_associate_spline_long_dev(_ptr(gps_t), _ptr(gps_xyz), M, _ptr(slam_t), N, float(gap), _ptr(work), _ptr(aligned), _ptr(valid),
                                           _ptr(status), _stream_ptr(stream))
    _lib.check(rc, "gsf_associate_spline_long_dev")
    return aligned, valid, status


def ekf_strict_batched(ts, pos, quat, z, offsets, params, init_pos, init_quat, params_per_traj=False, stream=None):
    """Literal step-by-step EKF recursion, one thread per trajectory (gsf_ekf_strict_batched_dev)."""
    lib = _lib.load()
    _require_cuda(ts, pos, quat, z, offsets, params, init_pos, init_quat)
    B = offsets.numel() - 1
    out_pos = torch.empty_like(pos)
    out_quat = torch.empty_like(quat)
    status = torch.empty((B,), dtype=torch.int32, device=ts.device)
    rc = lib.gsf_ekf_strict_batched_dev(_ptr(ts), _ptr(pos), _ptr(quat), _ptr(z), _ptr(offsets), B, _ptr(params),
                                        int(bool(params_per_traj)), _ptr(init_pos), _ptr(init_quat),
                                        _ptr(out_pos), _ptr(out_quat), _ptr(status), _stream_ptr(stream))
    _lib.check(rc, "gsf_ekf_strict_batched_dev")
    return out_pos, out_quat, status


def sim3_ransac(src, dst, samples, residual_threshold, min_inliers, stream=None):
    """compute_sim3_transform_robust with host-supplied sample indices (gsf_sim3_ransac_dev).
    src/dst [n,3] device fp64, samples [trials, m] device int32.
    Returns (R [3,3], t [3], s [1], inlier_mask [n] uint8, info [3] int32, status [1] int32); asynchronous."""
    lib = _lib.load()
    _require_cuda(src, dst, samples)
    n = src.shape[0]
    trials, m = samples.shape
    dev = src.device
    work = torch.empty((max(1, lib.gsf_sim3_ransac_work_doubles(int(trials), int(n))),), dtype=torch.float64, device=dev)
    mask = torch.empty((n,), dtype=torch.uint8, device=dev)
    R = torch.empty((3, 3), dtype=torch.float64, device=dev)
    t = torch.empty((3,), dtype=torch.float64, device=dev)
    s = torch.empty((1,), dtype=torch.float64, device=dev)
    info = torch.empty((3,), dtype=torch.int32, device=dev)
    status = torch.empty((1,), dtype=torch.int32, device=dev)
    samples = samples.to(torch.int32).contiguous()
    rc = lib.gsf_sim3_ransac_dev(_ptr(src), _ptr(dst), int(n), _ptr(samples), int(trials), int(m), float(residual_threshold),
                                 int(min_inliers), _ptr(work), _ptr(mask), _ptr(R), _ptr(t), _ptr(s), _ptr(info), _ptr(status),
                                 _stream_ptr(stream))
    _lib.check(rc, "gsf_sim3_ransac_dev")
    return R, t, s, mask, info, status


def umeyama_batched(src, dst, offsets, max_len, mask=None, stream=None):
    """compute_sim3_transform for B point-set pairs -> (R [B,3,3], t [B,3], s [B], status [B])."""
    lib = _lib.load()
    _require_cuda(src, dst, offsets, mask)
    B = offsets.numel() - 1
    dev = src.device
    work = torch.empty((max(1, lib.gsf_umeyama_work_doubles(B, int(max_len))),), dtype=torch.float64, device=dev)
    R = torch.empty((B, 3, 3), dtype=torch.float64, device=dev)
    t = torch.empty((B, 3), dtype=torch.float64, device=dev)
    s = torch.empty((B,), dtype=torch.float64, device=dev)
    status = torch.empty((B,), dtype=torch.int32, device=dev)
    rc = lib.gsf_sim3_umeyama_batched_dev(_ptr(src), _ptr(dst), _ptr(offsets), _ptr(mask), B, int(max_len), _ptr(work),
                                          _ptr(R), _ptr(t), _ptr(s), _ptr(status), _stream_ptr(stream))
    _lib.check(rc, "gsf_sim3_umeyama_batched_dev")
    return R, t, s, status


SIM3_STATS = 20        # GSF_SIM3_STATS: count, mean src, mean dst, centred H, sum |src_c|^2, padding


def sim3_partial_stats(src, dst, mask=None, stream=None):
    """One shard's contribution to compute_sim3_transform of a trajectory spread over several GPUs
    (gsf_sim3_partial_stats_dev) -> [SIM3_STATS] float64."""
    lib = _lib.load()
    _require_cuda(src, dst, mask)
    n = int(src.shape[0])
    dev = src.device
    off = torch.tensor([0, n], dtype=torch.int64, device=dev)
    work = torch.empty((max(1, lib.gsf_umeyama_work_doubles(1, n)),), dtype=torch.float64, device=dev)
    stats = torch.empty((SIM3_STATS,), dtype=torch.float64, device=dev)
    rc = lib.gsf_sim3_partial_stats_dev(_ptr(src), _ptr(dst), _ptr(off), _ptr(mask), n, _ptr(work), _ptr(stats), _stream_ptr(stream))
    _lib.check(rc, "gsf_sim3_partial_stats_dev")
    return stats


def sim3_from_partial_stats(stats, stream=None):
    """R [3,3], t [3], s [1], status [1] from the statistics of all shards, [shards, SIM3_STATS] in shard order
    (gsf_sim3_from_partial_stats_dev)."""
    lib = _lib.load()
    _require_cuda(stats)
    stats = stats.contiguous()
    dev = stats.device
    R = torch.empty((3, 3), dtype=torch.float64, device=dev)
    t = torch.empty((3,), dtype=torch.float64, device=dev)
    s = torch.empty((1,), dtype=torch.float64, device=dev)
    status = torch.empty((1,), dtype=torch.int32, device=dev)
    rc = lib.gsf_sim3_from_partial_stats_dev(_ptr(stats), int(stats.shape[0]), _ptr(R), _ptr(t), _ptr(s), _ptr(status), _stream_ptr(stream))
    _lib.check(rc, "gsf_sim3_from_partial_stats_dev")
    return R, t, s, status


def sim3_apply_batched(pos, quat, offsets, max_len, R, t, s, stream=None):
    """transform_trajectory for B trajectories -> (pos', quat', status)."""
    lib = _lib.load()
    _require_cuda(pos, quat, offsets, R, t, s)
    B = offsets.numel() - 1
    out_pos = torch.empty_like(pos)
    out_quat = torch.empty_like(quat)
    status = torch.zeros((B,), dtype=torch.int32, device=pos.device)
    rc = lib.gsf_sim3_apply_dev(_ptr(pos), _ptr(quat), _ptr(offsets), _ptr(R), _ptr(t), _ptr(s), B, int(max_len),
                                _ptr(out_pos), _ptr(out_quat), _ptr(status), _stream_ptr(stream))
    _lib.check(rc, "gsf_sim3_apply_dev")
    return out_pos, out_quat, status


def ate_nn_batched(traj, cand, ts, offsets, max_len, skip=5.0, stream=None):
    """Nearest-neighbour error statistics -> stats [B,4] = mean, median, RMSE, count."""
    lib = _lib.load()
    _require_cuda(traj, cand, ts, offsets)
    B = offsets.numel() - 1
    stats = torch.empty((B, 4), dtype=torch.float64, device=traj.device)
    nwork = lib.gsf_ate_work_doubles(int(ts.numel()), int(max_len))
    work = torch.empty((nwork,), dtype=torch.float64, device=traj.device) if nwork > 0 else None
    rc = lib.gsf_ate_nn_batched_dev(_ptr(traj), _ptr(cand), _ptr(ts), _ptr(offsets), B, int(max_len), float(skip),
                                    _ptr(work), _ptr(stats), _stream_ptr(stream))
    _lib.check(rc, "gsf_ate_nn_batched_dev")
    return stats


def utm_forward(lon, lat, zone: int, south: bool, stream=None):
    lib = _lib.load()
    _require_cuda(lon, lat)
    east, north = torch.empty_like(lon), torch.empty_like(lat)
    rc = lib.gsf_utm_forward_dev(_ptr(lon), _ptr(lat), lon.numel(), int(zone), int(bool(south)), _ptr(east), _ptr(north),
                                 _stream_ptr(stream))
    _lib.check(rc, "gsf_utm_forward_dev")
    return east, north


def utm_inverse(east, north, zone: int, south: bool, stream=None):
    lib = _lib.load()
    _require_cuda(east, north)
    lon, lat = torch.empty_like(east), torch.empty_like(north)
    rc = lib.gsf_utm_inverse_dev(_ptr(east), _ptr(north), east.numel(), int(zone), int(bool(south)), _ptr(lon), _ptr(lat),
                                 _stream_ptr(stream))
    _lib.check(rc, "gsf_utm_inverse_dev")
    return lon, lat


def geo_zone(lon, lat, stream=None):
    """auto_utm_projection on the device -> tensor [4] = mean lon, mean lat, zone, south."""
    lib = _lib.load()
    _require_cuda(lon, lat)
    part = torch.empty((2 * _lib.GEO_PARTS,), dtype=torch.float64, device=lon.device)
    out = torch.empty((4,), dtype=torch.float64, device=lon.device)
    rc = lib.gsf_geo_zone_dev(_ptr(lon), _ptr(lat), lon.numel(), _ptr(part), _ptr(out), _stream_ptr(stream))
    _lib.check(rc, "gsf_geo_zone_dev")
    return out


def associate_spline(gps_t, gps_xyz, gps_offsets, slam_t, slam_offsets, gap=5.0, stream=None):
    """dynamic_time_alignment's per-segment interpolation -> (aligned [n,3], valid [n] uint8)."""
    lib = _lib.load()
    _require_cuda(gps_t, gps_xyz, gps_offsets, slam_t, slam_offsets)
    B = gps_offsets.numel() - 1
    dev = gps_t.device
    work = torch.empty((max(1, 4 * gps_t.numel()),), dtype=torch.float64, device=dev)
    aligned = torch.empty((slam_t.numel(), 3), dtype=torch.float64, device=dev)
    valid = torch.empty((slam_t.numel(),), dtype=torch.uint8, device=dev)
    rc = lib.gsf_associate_spline_dev(_ptr(gps_t), _ptr(gps_xyz), _ptr(gps_offsets), _ptr(slam_t), _ptr(slam_offsets),
                                      B, float(gap), _ptr(work), _ptr(aligned), _ptr(valid), _stream_ptr(stream))
    _lib.check(rc, "gsf_associate_spline_dev")
    return aligned, valid


def synth_generate(batch: int, n: int, dt: float, speed: float, seed: int, first_traj: int = 0,
                   outage_prob: float = 0.0, outage_max_len: int = 0, device="cuda", out=None, stream=None):
    """Device-side synthetic batch -> (ts [B*n], pos [B*n,3], quat [B*n,4], z [B*n,3])."""
    lib = _lib.load()
    P = batch * n
    if out is None:
        ts = torch.empty((P,), dtype=torch.float64, device=device)
        pos = torch.empty((P, 3), dtype=torch.float64, device=device)
        quat = torch.empty((P, 4), dtype=torch.float64, device=device)
        z = torch.empty((P, 3), dtype=torch.float64, device=device)
    else:
        ts, pos, quat, z = out
    rc = lib.gsf_synth_generate_dev(_ptr(ts), _ptr(pos), _ptr(quat), _ptr(z), int(first_traj), int(batch), int(n),
                                    float(dt), float(speed), int(seed), float(outage_prob), int(outage_max_len),
                                    _stream_ptr(stream))
    _lib.check(rc, "gsf_synth_generate_dev")
    return ts, pos, quat, z


def fuse_batched_host(ts, pos, quat, z, offsets, max_len, params_blob, params_per_traj=False,
                      init_pos=None, init_quat=None, out_pos=None, out_quat=None, sim3_out=None, status=None):
    """Host-buffer entry point (gsf_fuse_batched_host): numpy / pinned CPU tensors in and out,
    H2D + kernel + D2H pipelined inside the library.  Blocking."""
    lib = _lib.load()

    def hp(a):
        if a is None:
            return None
        if isinstance(a, torch.Tensor):
            return ctypes.c_void_p(a.data_ptr())
        return ctypes.c_void_p(a.ctypes.data)

    B = len(offsets) - 1
    P = int(offsets[-1])
    if out_pos is None:
        out_pos = np.empty((P, 3))
    if out_quat is None:
        out_quat = np.empty((P, 4))
    if sim3_out is None:
        sim3_out = np.empty((B, 16))
    if status is None:
        status = np.empty((B,), dtype=np.int32)
    rc = lib.gsf_fuse_batched_host(hp(ts), hp(pos), hp(quat), hp(z), hp(offsets), B, int(max_len), hp(params_blob),
                                   int(bool(params_per_traj)), hp(init_pos), hp(init_quat),
                                   hp(out_pos), hp(out_quat), hp(sim3_out), hp(status))
    _lib.check(rc, "gsf_fuse_batched_host")
    return out_pos, out_quat, sim3_out, status


__all__ = [n for n in dir() if not n.startswith("_")]
