#!/bin/bash
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --trajectories 65536 --no-cpu-baseline --no-e2e"
$CMD > gpurun_out/p1_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/p1_launches.csv $CMD > gpurun_out/p1_ncu_launch.log 2>&1
$CMD > gpurun_out/p1_plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:fuse_traj -s 3 -c 1 -o gpurun_out/p1_fuse $CMD > gpurun_out/p1_ncu_full.log 2>&1
tail -2 gpurun_out/p1_plain.log | cut -c1-400
tail -5 gpurun_out/p1_ncu_full.log
ls -la gpurun_out/
