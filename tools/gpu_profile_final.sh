#!/bin/bash
# Round profile artefacts: launch list of the bench command + full capture of the fused kernel.
mkdir -p gpurun_out
BENCH="python bench.py --steps 3 --warmup 3 --no-cpu-baseline"
$BENCH > gpurun_out/pf_bench_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/pf_launches.csv $BENCH > gpurun_out/pf_ncu_launch.log 2>&1
tail -1 gpurun_out/pf_bench_plain.log | cut -c1-300
CMD="python bench.py --steps 2 --warmup 3 --trajectories 65536 --no-cpu-baseline --no-e2e"
$CMD > gpurun_out/pf_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:fuse_traj -s 3 -c 1 -o gpurun_out/pf_fuse $CMD > gpurun_out/pf_ncu_full.log 2>&1
tail -2 gpurun_out/pf_ncu_full.log
