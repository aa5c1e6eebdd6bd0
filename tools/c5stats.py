import sys, numpy as np, torch
sys.path.insert(0, ".")
from gps_optimize_slam_b200 import fusion, synth
from gps_optimize_slam_b200.config import pack_fuse_params
tr = synth.make_loop_trajectory(2026, n=4541)
k = 64
d = lambda a, dt=torch.float64: torch.from_numpy(np.ascontiguousarray(a)).to("cuda", dtype=dt)
stats, sim3, st = fusion.noise_grid(d(tr["ts"]), d(tr["pos"]), d(tr["quat"]), d(tr["gps"]), d(pack_fuse_params(), torch.uint8),
                                    d(np.logspace(-3, 1, k)), d(np.logspace(-3, 1, k)), d(np.logspace(-2, 1, k)))
s = stats.cpu().numpy()
print("mean-error percentiles [m]:", np.percentile(s[:, 0], [0, 1, 10, 25, 50, 75, 90, 99, 100]).round(4))
print("fraction mean < 0.25 m, < 0.5 m, < 1 m:", (s[:, 0] < 0.25).mean(), (s[:, 0] < 0.5).mean(), (s[:, 0] < 1).mean())
z = tr["gps"]; dd = np.linalg.norm(z[1:] - z[:-1], axis=1)
print("adjacent candidate spacing: min %.3f median %.3f" % (dd.min(), np.median(dd)), "extent", np.ptp(z[:, 0]), np.ptp(z[:, 1]))
