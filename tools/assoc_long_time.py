"""Times gsf_associate_spline_long_dev on one synthetic track (run under ncu --metrics gpu__time_duration.sum for the per-kernel split).
Usage: python tools/assoc_long_time.py [M] [reps]"""
import sys, time
import torch
sys.path.insert(0, ".")
from gps_optimize_slam_b200 import fusion

M = int(float(sys.argv[1])) if len(sys.argv) > 1 else 100_000_000
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
dev = torch.device("cuda:0")
gt = torch.arange(M, device=dev, dtype=torch.float64) * 0.1 + 0.01 * torch.rand(M, device=dev, dtype=torch.float64)
gy = torch.stack([8.0 * gt + 30 * torch.sin(gt / 7.0), 3.0 * gt + 50 * torch.cos(gt / 11.0), torch.sin(gt / 3.0)], dim=1).contiguous()
gy += 0.3 * torch.randn_like(gy)
st = gt + 0.037
for r in range(reps):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    a, v, status = fusion.associate_spline_long(gt, gy, st, 5.0)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
print(f"M = N = {M}: {dt * 1e3:.3f} ms, {M / dt:.3e} stamps/s, valid {int(v.sum())}, status {int(status[0])}")
