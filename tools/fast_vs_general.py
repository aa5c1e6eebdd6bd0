"""A/B check of the two fused kernels on the same device-generated batch (GSF_FUSE_IMPL toggles the
dispatch inside gsf_fuse_batched_dev).  Usage: python tools/fast_vs_general.py [B] [n] [outage_prob]"""
import os, sys, time
import torch
sys.path.insert(0, ".")
from gps_optimize_slam_b200 import fusion

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
n = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
outage = float(sys.argv[3]) if len(sys.argv) > 3 else 0.0
ts, pos, quat, z = fusion.synth_generate(B, n, 0.1, 10.0, seed=7, outage_prob=outage, outage_max_len=20)
off = fusion.equal_offsets(B, n)
prm = fusion.params_tensor()
res = {}
for impl in ("general", "fast"):
    os.environ["GSF_FUSE_IMPL"] = impl
    for rep in range(3):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        out = fusion.fuse_batched(ts, pos, quat, z, off, n, prm)
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
    res[impl] = [o.clone() for o in out]
    print(f"{impl:8s} B={B} n={n} {dt*1e3:8.3f} ms  {B*n*144/dt/1e9:8.1f} GB/s  nonzero status {(out[3] != 0).sum().item()}", flush=True)
g, f = res["general"], res["fast"]
print("max |pos diff|", (g[0] - f[0]).abs().max().item(), " max |quat diff|", (g[1] - f[1]).abs().max().item(),
      " max |sim3 diff| (R,t,s)", (g[2][:, :9] - f[2][:, :9]).abs().max().item(), (g[2][:, 9:12] - f[2][:, 9:12]).abs().max().item(),
      (g[2][:, 12] - f[2][:, 12]).abs().max().item(), " counts equal", bool((g[2][:, 13:] == f[2][:, 13:]).all()),
      " status equal", bool((g[3] == f[3]).all()), " nan", bool(torch.isnan(f[0]).any()))
