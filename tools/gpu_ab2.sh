#!/bin/bash
mkdir -p gpurun_out
{
for ct in 128 64; do echo "CT=$ct"; GSF_FAST_CT=$ct timeout 100 python tools/phase_timing_fast.py 65536 1000; GSF_FAST_CT=$ct timeout 200 python tools/fast_vs_general.py 65536 1000 2>&1 | tail -2; done
} > gpurun_out/ab2.log 2>&1
cat gpurun_out/ab2.log
CMD="python bench.py --steps 2 --warmup 3 --trajectories 65536 --no-cpu-baseline --no-e2e"
timeout 900 ncu --set full --import-source on --clock-control none -k regex:fuse_fast -s 3 -c 1 -o gpurun_out/ab2_fast -f $CMD > gpurun_out/ab2_ncu.log 2>&1; tail -2 gpurun_out/ab2_ncu.log
