import os, sys, torch
sys.path.insert(0, ".")
from gps_optimize_slam_b200 import fusion
for n in (1088, 1100, 1152, 1500):
    ts, pos, quat, z = fusion.synth_generate(2048, n, 0.1, 10.0, seed=3)
    off = fusion.equal_offsets(2048, n); prm = fusion.params_tensor()
    out = fusion.fuse_batched(ts, pos, quat, z, off, n, prm)
    os.environ["GSF_FUSE_IMPL"] = "general"
    ref = fusion.fuse_batched(ts, pos, quat, z, off, n, prm)
    del os.environ["GSF_FUSE_IMPL"]
    print(n, "status0", int((out[3] != 0).sum()), "max pos diff", float((out[0] - ref[0]).abs().max()), "quat", float((out[1] - ref[1]).abs().max()))
