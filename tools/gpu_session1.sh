#!/bin/bash
# First GPU session: smoke, parity tests, small + default bench.  Everything under timeouts.
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/s1_gpu.txt 2>&1
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/s1_smoke.log 2>&1; echo "smoke rc=$?" | tee -a gpurun_out/s1_summary.txt
timeout 900 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/s1_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/s1_summary.txt
tail -80 gpurun_out/s1_pytest.log | cut -c1-220
timeout 600 python bench.py --steps 5 --warmup 3 --trajectories 65536 --no-cpu-baseline > gpurun_out/s1_bench_small.log 2>&1; echo "bench small rc=$?" | tee -a gpurun_out/s1_summary.txt
tail -3 gpurun_out/s1_bench_small.log
timeout 600 python bench.py --steps 5 --warmup 3 --workload config2 --no-cpu-baseline > gpurun_out/s1_bench_c2.log 2>&1; echo "bench config2 rc=$?" | tee -a gpurun_out/s1_summary.txt
tail -3 gpurun_out/s1_bench_c2.log
