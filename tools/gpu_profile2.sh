#!/bin/bash
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --trajectories 65536 --no-cpu-baseline --no-e2e"
$CMD > gpurun_out/p2_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:fuse_traj -s 3 -c 1 -o gpurun_out/p2_fuse $CMD > gpurun_out/p2_ncu_full.log 2>&1
tail -2 gpurun_out/p2_ncu_full.log
