#!/bin/bash
# One GPU session: parity tests, then whatever measurement the round needs.  Usage: gpurun -- bash tools/gpu_session.sh [tag] [extra...]
tag=${1:-s}
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/${tag}_gpu.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${tag}_pytest.log
tail -5 gpurun_out/${tag}_pytest.log
timeout 600 python bench.py > gpurun_out/${tag}_bench_c3.json 2> gpurun_out/${tag}_bench_c3.err; echo "bench rc=$?"
tail -c 3000 gpurun_out/${tag}_bench_c3.json
timeout 300 python bench.py --workload config2 --no-cpu-baseline > gpurun_out/${tag}_bench_c2.json 2> gpurun_out/${tag}_bench_c2.err; echo "bench c2 rc=$?"
timeout 300 python bench.py --workload config5 --no-cpu-baseline > gpurun_out/${tag}_bench_c5.json 2> gpurun_out/${tag}_bench_c5.err; echo "bench c5 rc=$?"
tail -c 1500 gpurun_out/${tag}_bench_c5.json
