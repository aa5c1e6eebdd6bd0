#!/bin/bash
bash tools/gpu_variants.sh np1 > /dev/null 2>&1
grep -E "^==|^fast" gpurun_out/variants.log
cp variants/libgsf_base.so gps_optimize_slam_b200/libgsf.so
for ct in 64 128; do echo "== base CT=$ct"; GSF_FAST_CT=$ct timeout 200 python tools/fast_vs_general.py 65536 1000 2>&1 | grep -E "^fast"; done
timeout 600 python -m pytest tests -m gpu -x -q -k "fus or fast or batched" 2>&1 | tail -2
