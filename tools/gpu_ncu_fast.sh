#!/bin/bash
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --trajectories 65536 --no-cpu-baseline --no-e2e"
timeout 300 $CMD > gpurun_out/nf_plain.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fuse_fast -s 3 -c 1 -f -o gpurun_out/nf_fast $CMD > gpurun_out/nf_ncu.log 2>&1
tail -2 gpurun_out/nf_ncu.log
