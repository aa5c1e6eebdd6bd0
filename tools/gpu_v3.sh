#!/bin/bash
mkdir -p gpurun_out
for args in "512 271" "2048 1000" "1024 1000 0.02" "3000 500"; do timeout 120 python tools/fast_vs_general.py $args 2>&1 | tail -4; echo "rc=$?"; done > gpurun_out/v3_ab.log 2>&1
cat gpurun_out/v3_ab.log
