"""Debug: per-role cycle stamps of one trajectory (the 5th of block 0) inside fuse_fast_kernel."""
import ctypes, sys, torch
sys.path.insert(0, ".")
from gps_optimize_slam_b200 import _lib, fusion
lib = _lib.load()
B, n = int(sys.argv[1]), int(sys.argv[2])
ts, pos, quat, z = fusion.synth_generate(B, n, 0.1, 10.0, seed=1)
off = fusion.equal_offsets(B, n); prm = fusion.params_tensor()
buf = torch.zeros(128, dtype=torch.int64, device="cuda")
lib.gsf_debug_phase_clock.argtypes = [ctypes.c_void_p]; lib.gsf_debug_phase_clock.restype = None
lib.gsf_debug_phase_clock(ctypes.c_void_p(buf.data_ptr()))
for _ in range(3):
    fusion.fuse_batched(ts, pos, quat, z, off, n, prm)
torch.cuda.synchronize()
c = buf.cpu().tolist()
base = min(x for x in c if x > 0)
def rel(k): return c[k] - base
print("compute (traj j=3): start %d | load landed +%d | aux ready +%d | passB+scan +%d | passC +%d | quats +%d | end +%d  (period end-start %d)" % (
    rel(0), c[1]-c[0], c[2]-c[1], c[3]-c[2], c[4]-c[3], c[5]-c[4], c[6]-c[5], c[6]-c[0]))
print("warp A  (traj j=3): start %d | wait free +%d | stream sums +%d | butterfly+publish +%d" % (rel(16), c[17]-c[16], c[18]-c[17], c[19]-c[18]))
print("warp B  (traj j=3): start %d | wait free +%d | cov scan +%d | wait sums +%d | SVD+publish +%d" % (rel(24), c[25]-c[24], c[26]-c[25], c[27]-c[26], c[28]-c[27]))
print("warp B scan detail: ts wait %d | step loops %d | warp scan %d | eval+publish %d" % (c[31]-c[25], c[29]-c[31], c[30]-c[29], c[26]-c[30]))
print('per-block totals (block, sm, trajectories, cycles/trajectory):', [(37*k, c[64+4*k+2], c[64+4*k+1], c[64+4*k]//max(1,c[64+4*k+1])) for k in range(12)])
lib.gsf_debug_phase_clock(None)
