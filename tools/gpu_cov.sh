#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/q_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/q_pytest.log | cut -c1-300
timeout 600 python bench.py --workload config5 --steps 2 --warmup 3 --grid-k 16 > gpurun_out/c5_small.log 2>&1; tail -1 gpurun_out/c5_small.log | cut -c1-1800
timeout 900 python bench.py --workload config5 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/c5_full.log 2>&1; tail -1 gpurun_out/c5_full.log | cut -c1-1800
timeout 900 python bench.py --workload config4 --steps 3 --warmup 3 > gpurun_out/c4_full.log 2>&1; tail -1 gpurun_out/c4_full.log | cut -c1-2200
