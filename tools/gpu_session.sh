#!/bin/bash
# One GPU session: parity tests, then whatever measurement the round needs.  Usage: gpurun -- bash tools/gpu_session.sh [tag]
tag=${1:-s}
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/${tag}_gpu.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${tag}_pytest.log
tail -5 gpurun_out/${tag}_pytest.log
