"""Times gsf_ate_nn_batched_dev on fused synthetic trajectories.  Usage: [GSF_LIB=...] python tools/ate_time.py [B] [n] [reps]"""
import statistics, sys
import torch
sys.path.insert(0, ".")
from gps_optimize_slam_b200 import fusion
B = int(sys.argv[1]) if len(sys.argv) > 1 else 262144
n = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 5
ts, pos, quat, z = fusion.synth_generate(B, n, 0.1, 10.0, seed=7)
off = fusion.equal_offsets(B, n); prm = fusion.params_tensor()
out = fusion.fuse_batched(ts, pos, quat, z, off, n, prm)
times = []
for r in range(reps + 2):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); stats = fusion.ate_nn_batched(out[0], z, ts, off, n); e1.record()
    torch.cuda.synchronize()
    if r >= 2: times.append(e0.elapsed_time(e1))
print("B=%d n=%d  ATE min %.3f ms median %.3f ms  -> %.1f GB/s  checksum %.17g %.17g" % (
    B, n, min(times), statistics.median(times), B * n * 56 / min(times) / 1e6, float(stats[:, :3].sum()), float(stats[:, 1].sum())))
