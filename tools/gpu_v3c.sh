#!/bin/bash
mkdir -p gpurun_out
{
for ct in 96 128; do echo "CT=$ct"; GSF_FAST_CT=$ct timeout 100 python tools/phase_timing_fast.py 65536 1000; GSF_FAST_CT=$ct timeout 200 python tools/fast_vs_general.py 65536 1000 2>&1 | tail -3; done
timeout 200 python tools/fast_vs_general.py 65536 271 | tail -3
} > gpurun_out/v3c.log 2>&1
cat gpurun_out/v3c.log
CMD="python bench.py --steps 2 --warmup 3 --trajectories 65536 --no-cpu-baseline --no-e2e"
timeout 300 $CMD > gpurun_out/v3c_plain.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fuse_fast -s 3 -c 1 -o gpurun_out/v3c_fast $CMD > gpurun_out/v3c_ncu.log 2>&1
tail -2 gpurun_out/v3c_ncu.log
