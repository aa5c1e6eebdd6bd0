#!/bin/bash
bash tools/gpu_variants.sh "$@" > /dev/null 2>&1
grep -E "^==|^fast" gpurun_out/variants.log | cut -c1-200
last="${@: -1}"
cp variants/libgsf_$last.so gps_optimize_slam_b200/libgsf.so
timeout 600 python -m pytest tests -m gpu -x -q -k "fus or fast or batched" 2>&1 | tail -2
CMD="python bench.py --steps 2 --warmup 3 --trajectories 65536 --no-cpu-baseline --no-e2e"
timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:fuse_fast -s 3 -c 1 $CMD 2>&1 | grep -E "dram__|gpu__time"
