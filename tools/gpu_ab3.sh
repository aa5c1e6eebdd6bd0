#!/bin/bash
# A/B + parity tests of the fused path, then one ncu --set full capture of the fast kernel
bash tools/gpu_ab.sh
CMD="python bench.py --steps 2 --warmup 3 --trajectories 65536 --no-cpu-baseline --no-e2e"
timeout 900 ncu --set full --import-source on --clock-control none -k regex:fuse_fast -s 3 -c 1 -o gpurun_out/ab3_fast -f $CMD > gpurun_out/ab3_ncu.log 2>&1; tail -1 gpurun_out/ab3_ncu.log
