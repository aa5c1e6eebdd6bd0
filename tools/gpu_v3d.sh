#!/bin/bash
mkdir -p gpurun_out
{
for ct in 96; do echo "CT=$ct"; GSF_FAST_CT=$ct timeout 100 python tools/phase_timing_fast.py 65536 1000; GSF_FAST_CT=$ct timeout 200 python tools/fast_vs_general.py 65536 1000 2>&1 | tail -3; done
timeout 200 python tools/fast_vs_general.py 65536 271 | tail -3
} > gpurun_out/v3d.log 2>&1
cat gpurun_out/v3d.log
CMD="python bench.py --steps 2 --warmup 3 --trajectories 65536 --no-cpu-baseline --no-e2e"
timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum --clock-control none -k regex:fuse_fast -s 3 -c 1 $CMD 2>&1 | grep -E "dram__|gpu__time|lts__|fp64|issue_active|inst_executed"
