"""Short trajectories: the one-warp-per-trajectory kernel against the role kernel <32, 9> (GSF_FAST_CT=32) on the same batch:
timing and bit-equality.  Usage: python tools/warp_vs_roles.py [B] [n] [outage_prob]"""
import os, statistics, sys, time
import torch
sys.path.insert(0, ".")
from gps_optimize_slam_b200 import fusion
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
n = int(sys.argv[2]) if len(sys.argv) > 2 else 271
outage = float(sys.argv[3]) if len(sys.argv) > 3 else 0.0
ts, pos, quat, z = fusion.synth_generate(B, n, 0.104, 13.0, seed=7, outage_prob=outage, outage_max_len=20)
off = fusion.equal_offsets(B, n); prm = fusion.params_tensor()
res = {}
for name, env in (("roles", "32"), ("warp", "")):
    if env: os.environ["GSF_FAST_CT"] = env
    else: os.environ.pop("GSF_FAST_CT", None)
    times = []
    for r in range(9):
        torch.cuda.synchronize(); time.sleep(0.02)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); out = fusion.fuse_batched(ts, pos, quat, z, off, n, prm); e1.record()
        torch.cuda.synchronize()
        if r >= 2: times.append(e0.elapsed_time(e1))
    res[name] = [o.clone() for o in out]
    print("%-6s B=%d n=%d  min %.4f ms  median %.4f ms -> %.1f GB/s  nonzero status %d" % (name, B, n, min(times), statistics.median(times), B * n * 144 / min(times) / 1e6, int((out[3] != 0).sum())))
a, b = res["roles"], res["warp"]
print("bit-equal: pos", bool(torch.equal(a[0], b[0])), "quat", bool(torch.equal(a[1], b[1])), "sim3", bool(torch.equal(a[2], b[2])), "status", bool(torch.equal(a[3], b[3])),
      " nan", bool(torch.isnan(b[0]).any()))
