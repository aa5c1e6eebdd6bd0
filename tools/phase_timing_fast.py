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
nt = max(1, c[65]) * 3          # trajectories of block 0 over the three calls
a = lambda k: c[k] / nt
print("block 0, averages over %d trajectories (cycles)" % nt)
print("compute: queue %.0f | load wait %.0f | aux wait %.0f | passB+scan %.0f | passC %.0f | quats %.0f | end %.0f | period %.0f" % (
    a(0), a(1), a(2), a(3), a(4), a(5), a(6), sum(a(k) for k in range(7))))
print("warp A : top %.0f | wait start %.0f | stream sums %.0f | butterfly+wait slot+publish %.0f | period %.0f" % (a(16), a(17), a(18), a(19), sum(a(k) for k in range(16, 20))))
print("warp B : top %.0f | wait free %.0f | cov scan %.0f | wait sums %.0f | SVD+publish %.0f | period %.0f" % (a(24), a(25), a(26), a(27), a(28), sum(a(k) for k in range(24, 29))))
print('per-block totals (block, sm, trajectories, cycles/trajectory):', [(37*k, c[64+4*k+2], c[64+4*k+1], c[64+4*k]//max(1,c[64+4*k+1])) for k in range(12)])
lib.gsf_debug_phase_clock(None)
