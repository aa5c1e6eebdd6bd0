#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/q_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/q_pytest.log | cut -c1-300
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_full.log 2>&1; tail -1 gpurun_out/bench_full.log | cut -c1-1500
timeout 600 python bench.py --steps 5 --warmup 3 --workload config2 --no-cpu-baseline > gpurun_out/bench_c2.log 2>&1; tail -1 gpurun_out/bench_c2.log | cut -c1-600
