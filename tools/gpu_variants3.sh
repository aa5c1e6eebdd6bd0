#!/bin/bash
bash tools/gpu_variants.sh "$@" > /dev/null 2>&1
grep -E "^==|^fast|^max" gpurun_out/variants.log | cut -c1-200
last="${@: -1}"
cp variants/libgsf_$last.so gps_optimize_slam_b200/libgsf.so
timeout 600 python -m pytest tests -m gpu -x -q -k "fus or fast or batched" 2>&1 | tail -2
