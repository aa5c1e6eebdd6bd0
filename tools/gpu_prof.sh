#!/bin/bash
# one ncu --set full capture of the fast kernel (65536 x 1000) after a plain run of the same command
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --trajectories 65536 --no-cpu-baseline --no-e2e"
timeout 300 $CMD > gpurun_out/prof_plain.log 2>&1; echo "plain rc=$?"; tail -1 gpurun_out/prof_plain.log | cut -c1-200
timeout 900 ncu --set full --import-source on --clock-control none -k regex:fuse_fast -s 3 -c 1 -o gpurun_out/prof_fast -f $CMD > gpurun_out/prof_ncu.log 2>&1; tail -1 gpurun_out/prof_ncu.log
