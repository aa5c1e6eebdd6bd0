#!/bin/bash
# time prebuilt kernel variants (variants/libgsf_<name>.so) on the same batch; GSF_FUSE_IMPL=fast only
mkdir -p gpurun_out
{
for v in base "$@" base; do
  cp variants/libgsf_$v.so gps_optimize_slam_b200/libgsf.so
  echo "== $v"
  timeout 200 python tools/fast_vs_general.py 65536 1000 2>&1 | grep -E "^fast|^max" | cut -c1-170
  timeout 200 python tools/fast_vs_general.py 65536 271 2>&1 | grep -E "^fast" | cut -c1-170
done
} > gpurun_out/variants.log 2>&1
cat gpurun_out/variants.log
