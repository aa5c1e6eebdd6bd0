"""Debug: per-phase cycle counts of one trajectory inside fuse_traj_kernel (block 0)."""
import ctypes, sys, torch
sys.path.insert(0, ".")
from gps_optimize_slam_b200 import _lib, fusion
lib = _lib.load()
B, n = int(sys.argv[1]), int(sys.argv[2])
outage = float(sys.argv[3]) if len(sys.argv) > 3 else 0.0            # share of trajectories with GNSS outages
ts, pos, quat, z = fusion.synth_generate(B, n, 0.1, 10.0, seed=1, outage_prob=outage, outage_max_len=20)
off = fusion.equal_offsets(B, n); prm = fusion.params_tensor()
buf = torch.zeros(16, dtype=torch.int64, device="cuda")
lib.gsf_debug_phase_clock.argtypes = [ctypes.c_void_p]; lib.gsf_debug_phase_clock.restype = None
lib.gsf_debug_phase_clock(ctypes.c_void_p(buf.data_ptr()))
for _ in range(3):
    fusion.fuse_batched(ts, pos, quat, z, off, n, prm)
torch.cuda.synchronize()
c = buf.cpu().tolist()
names = ["load-wait", "pass1 flags+sums", "general-sel + pass2 moebius + SVD", "pass3 gains", "pass4 final+rts", "store+next load+quat"]
print("n", n, "total cycles", c[6] - c[0])
for k, nm in enumerate(names):
    print(f"  {nm:36s} {c[k+1]-c[k]:8d}")
print("  [pass2 detail] moebius+scan", c[7]-c[2], " sums-combine", c[8]-c[7], " svd call", c[9]-c[8], " bcast+sync", c[3]-c[9])
if c[10] and c[11]:
    print("  [2->7 detail] sharp-turn gate", c[10]-c[2], " recovery flags", c[11]-c[10], " Sim3 selection + centred sums", c[7]-c[11])
lib.gsf_debug_phase_clock(None)
