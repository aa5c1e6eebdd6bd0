"""Multi-GPU plumbing.  Batched path: independent trajectories (or noise hypotheses) are sharded by
contiguous ranges, one process per GPU; there is no collective inside the data path, the only
exchange is the final gather of per-trajectory ATE statistics.  One long trajectory (config 4): the
poses are cut into contiguous blocks and three small all-gathers carry what crosses the cuts (zone
means, spline halo knots, Umeyama statistics).  torch.distributed: NCCL over NVLink on the GPUs,
gloo in the CPU tests."""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_range(total: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous, balanced split of [0, total): rank r owns [r*total//world, (r+1)*total//world)."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError(f"bad rank/world: {rank}/{world}")
    return rank * total // world, (rank + 1) * total // world


def gather_stats(local: torch.Tensor, total_rows: int | None = None, counts: list[int] | None = None) -> torch.Tensor:
    """All-gather row blocks of possibly different heights (shards differ by at most one row) and
    return them concatenated in rank order on every rank.  No-op without a process group.
    ``counts`` (rows per rank, e.g. from shard_range) saves the exchange of the block heights and every
    host synchronisation: the call is then one collective, asynchronous on the current stream."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return local
    world = dist.get_world_size()
    if counts is None:
        rows = torch.tensor([local.shape[0]], dtype=torch.int64, device=local.device)
        all_rows = [torch.zeros_like(rows) for _ in range(world)]
        dist.all_gather(all_rows, rows)
        counts = [int(r.item()) for r in all_rows]
    width = max(counts)
    if local.shape[0] == width:
        padded = local.contiguous()
    else:
        padded = torch.zeros((width,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        padded[: local.shape[0]] = local
    flat = torch.empty((world * width,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(flat, padded)
    if all(c == width for c in counts):
        out = flat
    else:
        out = torch.cat([flat[r * width: r * width + c] for r, c in enumerate(counts)])
    if total_rows is not None and out.shape[0] != total_rows:
        raise RuntimeError(f"gathered {out.shape[0]} rows, expected {total_rows}")
    return out


def max_over_ranks(value: float, device) -> float:
    """Timing reduction: the slowest rank defines the step time."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


# ----------------------------------------------------------------------------- one long trajectory over several GPUs
# BASELINE config 4 sharded: contiguous blocks of poses per rank.  Unlike the batched path this one HAS exchange
# steps, all of them tiny: the zone needs the mean longitude / latitude of the whole track (5 doubles per rank), the spline
# association needs the neighbours' first / last knots (the local-halo solve of gsf_assoc_long.cu makes 22 knots enough),
# the Umeyama fit needs every shard's statistics (20 doubles per rank).  Each is one all-gather; nothing else crosses ranks.
ASSOC_HALO = 22            # knots of the neighbouring shards the spline solve needs (AL_H + 2 of csrc/gsf_assoc_long.cu)


def _world():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def all_gather_rows(local: torch.Tensor) -> torch.Tensor:
    """[...] on every rank -> [world, ...] in rank order (equal shapes; one all_gather_into_tensor)."""
    rank, world = _world()
    if world == 1:
        return local.unsqueeze(0)
    flat = local.contiguous().reshape(-1)
    out = torch.empty((world * flat.numel(),), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, flat)
    return out.reshape((world,) + tuple(local.shape))


def global_zone(zone_local: torch.Tensor):
    """zone [5] = mean lon, mean lat, zone, south, valid count of THIS rank's rows (gsf_gnss_rows_to_utm_dev) ->
    (zone, south) of the whole track (auto_utm_projection, EKFGPSSLAM.py:127-134, from the count-weighted means), whether
    this rank's projection used them, and the valid-row counts of all ranks (list).  One all-gather of 5 doubles per rank; the
    result comes back to the host (one synchronisation)."""
    z = all_gather_rows(zone_local).cpu()
    cnt = z[:, 4]
    tot = float(cnt.sum())
    if tot <= 0:
        raise ValueError("no valid GNSS row on any rank")
    w = cnt > 0
    mean_lon = float((z[w, 0] * cnt[w]).sum() / tot)
    mean_lat = float((z[w, 1] * cnt[w]).sum() / tot)
    zone = int((mean_lon + 180.0) // 6 + 1)
    south = mean_lat < 0.0
    rank, _ = _world()
    mine_ok = (not bool(w[rank])) or (int(z[rank, 2]) == zone and bool(z[rank, 3] != 0) == south)
    return zone, south, mine_ok, [int(c) for c in cnt]


def exchange_halo(t_local: torch.Tensor, xyz_local: torch.Tensor, halo: int = ASSOC_HALO):
    """Knots of the neighbouring shards: returns (t_ext, xyz_ext, n_left) = this rank's knots with up to `halo` knots of the
    previous rank in front and of the next rank behind (none at the ends of the track).  One all-gather of [2, halo, 4]
    doubles per rank; shards shorter than `halo` contribute what they have (NaN-padded rows are dropped)."""
    rank, world = _world()
    if world == 1:
        return t_local, xyz_local, 0
    n = int(t_local.shape[0])
    edge = torch.full((2, halo, 4), float("nan"), dtype=torch.float64, device=t_local.device)
    k = min(halo, n)
    if k > 0:
        edge[0, :k, 0] = t_local[:k]; edge[0, :k, 1:] = xyz_local[:k]               # my first knots (for the previous rank)
        edge[1, halo - k:, 0] = t_local[n - k:]; edge[1, halo - k:, 1:] = xyz_local[n - k:]   # my last knots (for the next rank)
    alle = all_gather_rows(edge)
    parts_t, parts_x, n_left = [], [], 0
    if rank > 0:
        left = alle[rank - 1, 1]
        left = left[~torch.isnan(left[:, 0])]
        parts_t.append(left[:, 0]); parts_x.append(left[:, 1:]); n_left = int(left.shape[0])
    parts_t.append(t_local); parts_x.append(xyz_local)
    if rank < world - 1:
        right = alle[rank + 1, 0]
        right = right[~torch.isnan(right[:, 0])]
        parts_t.append(right[:, 0]); parts_x.append(right[:, 1:])
    return torch.cat(parts_t).contiguous(), torch.cat(parts_x).contiguous(), n_left


def long_trajectory_sharded(rows_local, slam_ts_local, slam_pos_local, slam_quat_local, gap: float = 5.0):
    """GNSS ingest -> spline association -> Sim3 (Umeyama over the whole track) -> transform, for ONE trajectory whose
    poses are split into contiguous blocks over the ranks.  rows_local [m,4] = ts, lat, lon, alt (GNSS samples of this
    block, sorted; at least 22 valid ones per rank), slam_* the SLAM poses of this block.  Returns (aligned [n,3], valid [n], R, t, s, out_pos, out_quat,
    (zone, south), association status, Sim3 status); R, t, s are identical on every rank."""
    from . import fusion
    rank, world = _world()
    dev = rows_local.device
    m, h = int(rows_local.shape[0]), ASSOC_HALO
    # 1. ingest with the zone of the whole track, written into buffers that have room for the neighbours' knots on both sides
    t_buf = torch.empty((m + 2 * h,), dtype=torch.float64, device=dev)
    x_buf = torch.empty((m + 2 * h, 3), dtype=torch.float64, device=dev)
    g_ts, g_xyz = t_buf[h:h + m], x_buf[h:h + m]
    _, _, zone_local = fusion.gnss_rows_to_utm(rows_local, out_ts=g_ts, out_xyz=g_xyz)
    zone, south, mine_ok, counts = global_zone(zone_local)
    if world > 1 and min(counts) < h:                       # the halo comes from the direct neighbours only
        raise ValueError(f"every rank needs at least {h} valid GNSS samples of the trajectory (got {counts})")
    if not mine_ok:                                         # this block's own means point to another zone: project again
        e, nn = fusion.utm_forward(rows_local[:, 2].contiguous(), rows_local[:, 1].contiguous(), zone, south)
        bad = torch.isnan(g_xyz[:, 0])
        g_xyz[:, 0] = e; g_xyz[:, 1] = nn
        g_xyz[bad] = float("nan")
    if counts[rank] != m:                                   # invalid rows: the spline takes the valid knots only
        keep = ~torch.isnan(g_xyz[:, 0])
        kt, kx = g_ts[keep], g_xyz[keep]
        m = int(kt.shape[0])
        t_buf = torch.empty((m + 2 * h,), dtype=torch.float64, device=dev); x_buf = torch.empty((m + 2 * h, 3), dtype=torch.float64, device=dev)
        t_buf[h:h + m] = kt; x_buf[h:h + m] = kx
        g_ts, g_xyz = t_buf[h:h + m], x_buf[h:h + m]
    # 2. association: the neighbours' edge knots (one all-gather) go into the margins -- no copy of the block itself
    k_l = min(h, counts[rank - 1]) if rank > 0 else 0
    k_r = min(h, counts[rank + 1]) if rank < world - 1 else 0
    if world > 1:
        edge = torch.full((2, h, 4), float("nan"), dtype=torch.float64, device=dev)
        k = min(h, m)
        if k > 0:
            edge[0, :k, 0] = g_ts[:k]; edge[0, :k, 1:] = g_xyz[:k]
            edge[1, h - k:, 0] = g_ts[m - k:]; edge[1, h - k:, 1:] = g_xyz[m - k:]
        alle = all_gather_rows(edge)
        if k_l:
            t_buf[h - k_l:h] = alle[rank - 1, 1, h - k_l:, 0]; x_buf[h - k_l:h] = alle[rank - 1, 1, h - k_l:, 1:]
        if k_r:
            t_buf[h + m:h + m + k_r] = alle[rank + 1, 0, :k_r, 0]; x_buf[h + m:h + m + k_r] = alle[rank + 1, 0, :k_r, 1:]
    aligned, valid, status = fusion.associate_spline_long(t_buf[h - k_l:h + m + k_r], x_buf[h - k_l:h + m + k_r], slam_ts_local, gap)
    # 3. Umeyama: every shard's statistics, merged in rank order on every rank
    stats = all_gather_rows(fusion.sim3_partial_stats(slam_pos_local, aligned, mask=valid))
    R, t, s, st = fusion.sim3_from_partial_stats(stats)
    # 4. transform of this block
    n = int(slam_pos_local.shape[0])
    off = torch.tensor([0, n], dtype=torch.int64, device=dev)
    out_pos, out_quat, _ = fusion.sim3_apply_batched(slam_pos_local, slam_quat_local, off, n, R.reshape(1, 3, 3), t.reshape(1, 3), s)
    return aligned, valid, R, t, s, out_pos, out_quat, (zone, south), status, st
