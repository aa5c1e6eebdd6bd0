"""One long trajectory over several GPUs (sharding.long_trajectory_sharded) against the single-GPU path on the same data.
Run under torchrun:  python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/config4_sharded_check.py [poses]"""
import os, sys
import torch
import torch.distributed as dist
sys.path.insert(0, ".")
from gps_optimize_slam_b200 import fusion, sharding

n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 2_000_000
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
g = torch.Generator(device=dev); g.manual_seed(7)                       # the same track on every rank
t = torch.arange(n, device=dev, dtype=torch.float64) * 0.1
lat = 49.0 + 1e-6 * torch.cumsum(torch.randn(n, device=dev, dtype=torch.float64, generator=g), 0)
lon = 11.97 + 2e-6 * torch.cumsum(torch.randn(n, device=dev, dtype=torch.float64, generator=g).abs(), 0)     # drifts across the 12 deg zone boundary
rows = torch.stack([t + 0.037, lat, lon, torch.full_like(t, 120.0)], dim=1).contiguous()
rows[n // 3: n // 3 + 200, 1] = 0.0                                     # invalid rows (a GNSS gap of 20 s) and
rows[5, 2] = 0.0
pos = torch.cumsum(torch.randn(n, 3, device=dev, dtype=torch.float64, generator=g), 0)
quat = torch.nn.functional.normalize(torch.randn(n, 4, device=dev, dtype=torch.float64, generator=g), dim=1)

# single-GPU reference on the whole track
g_ts, g_xyz, zone = fusion.gnss_rows_to_utm(rows)
keep = ~torch.isnan(g_xyz[:, 0])
a_ref, v_ref, _ = fusion.associate_spline_long(g_ts[keep].contiguous(), g_xyz[keep].contiguous(), t, 5.0)
off = torch.tensor([0, n], dtype=torch.int64, device=dev)
R_ref, t_ref, s_ref, _ = fusion.umeyama_batched(pos, a_ref, off, n, mask=v_ref)
p_ref, q_ref, _ = fusion.sim3_apply_batched(pos, quat, off, n, R_ref, t_ref, s_ref)

lo, hi = sharding.shard_range(n, rank, world)
aligned, valid, R, tt, s, out_pos, out_quat, zs, st_a, st_s = sharding.long_trajectory_sharded(rows[lo:hi].contiguous(), t[lo:hi].contiguous(),
                                                                                             pos[lo:hi].contiguous(), quat[lo:hi].contiguous())
z = zone.cpu()
ok_zone = zs == (int(z[2]), bool(z[3] != 0))
v_equal = bool(torch.equal(valid, v_ref[lo:hi]))
m = valid.bool()
d_al = float((aligned[m] - a_ref[lo:hi][m]).abs().max())
d_R = float((R - R_ref[0]).abs().max()); d_s = float((s - s_ref).abs().max()); d_t = float((tt - t_ref[0]).abs().max())
d_p = float((out_pos - p_ref[lo:hi]).abs().max()); d_q = float((out_quat - q_ref[lo:hi]).abs().max())
print(f"rank {rank}/{world}: zone {zs} ok {ok_zone}  valid equal {v_equal} ({int(m.sum())} of {hi - lo})  max |aligned diff| {d_al:.3e}  "
      f"|R| {d_R:.2e} |s| {d_s:.2e} |t| {d_t:.2e}  |pos| {d_p:.3e} |quat| {d_q:.2e}  status {int(st_a[0])} {int(st_s[0])}", flush=True)
assert ok_zone and v_equal and d_al < 1e-8 and d_R < 1e-12 and d_s < 1e-12 and d_t < 1e-6 and d_p < 1e-5 and d_q < 1e-12
if world > 1:
    dist.barrier(); dist.destroy_process_group()
