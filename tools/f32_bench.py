"""fp32 mode vs fp64 fused kernel on the same device-generated batch.  Usage: python tools/f32_bench.py [B] [n]"""
import sys, time
import torch
sys.path.insert(0, ".")
from gps_optimize_slam_b200 import fusion
B = int(sys.argv[1]) if len(sys.argv) > 1 else 262144
n = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
ts, pos, quat, z = fusion.synth_generate(B, n, 0.1, 10.0, seed=7)
off = fusion.equal_offsets(B, n); prm = fusion.params_tensor()
a = fusion.to_local_f32(ts, pos, quat, z, off)
op = torch.empty_like(a.pos32); oq = torch.empty_like(a.quat32)
s3 = torch.empty((B, 16), dtype=torch.float64, device="cuda"); st = torch.empty((B,), dtype=torch.int32, device="cuda")
for r in range(3):
    fusion.fuse_batched_f32(a, prm, out_pos=op, out_quat=oq, sim3_out=s3, status=st)
torch.cuda.synchronize()
ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
ev[0].record()
for r in range(5):
    fusion.fuse_batched_f32(a, prm, out_pos=op, out_quat=oq, sim3_out=s3, status=st)
ev[1].record(); torch.cuda.synchronize()
ms = ev[0].elapsed_time(ev[1]) / 5
print("fp32 mode: B=%d n=%d  %.3f ms  %.3e pose-updates/s  %.0f GB/s (72 B/pose)  bad %d" % (B, n, ms, B * (n - 1) / ms * 1e3, B * n * 72 / ms / 1e6, int((st != 0).sum())))
p64, q64, s64, st64 = fusion.fuse_batched(ts, pos, quat, z, off, n, prm)
p32, q32 = fusion.from_local_f32(a, op, oq)
print("max |fp32 mode - fp64 kernel| = %.3e m, quat %.2e" % ((p32 - p64).abs().max().item(), (q32 - q64).abs().max().item()))
