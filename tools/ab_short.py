"""Burst timing of gsf_fuse_batched_dev (default 65 536 x 1000 poses: ~2 ms, too short for the power cap to bite), CUDA events,
with idle gaps between launches.  Usage: GSF_LIB=... python tools/ab_short.py [B] [n] [reps]"""
import statistics, sys, time
import torch
sys.path.insert(0, ".")
from gps_optimize_slam_b200 import fusion
B = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
n = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 9
ts, pos, quat, z = fusion.synth_generate(B, n, 0.1, 10.0, seed=7)
off = fusion.equal_offsets(B, n); prm = fusion.params_tensor()
op = torch.empty_like(pos); oq = torch.empty_like(quat)
s3 = torch.empty((B, 16), dtype=torch.float64, device="cuda"); st = torch.empty((B,), dtype=torch.int32, device="cuda")
times = []
for r in range(reps + 2):
    torch.cuda.synchronize(); time.sleep(0.05)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); fusion.fuse_batched(ts, pos, quat, z, off, n, prm, out_pos=op, out_quat=oq, sim3_out=s3, status=st); e1.record()
    torch.cuda.synchronize()
    if r >= 2: times.append(e0.elapsed_time(e1))
print("B=%d n=%d  min %.4f ms  median %.4f ms  -> %.1f GB/s (min)  bad %d" % (B, n, min(times), statistics.median(times), B * n * 144 / min(times) / 1e6, int((st != 0).sum())))
