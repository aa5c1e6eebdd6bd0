#!/bin/bash
mkdir -p gpurun_out
for ct in 96 128; do echo "GSF_FAST_CT=$ct"; GSF_FAST_CT=$ct timeout 200 python tools/fast_vs_general.py 65536 1000 2>&1 | tail -3; done > gpurun_out/v3_scale.log 2>&1
timeout 200 python tools/fast_vs_general.py 65536 271 >> gpurun_out/v3_scale.log 2>&1
cat gpurun_out/v3_scale.log
timeout 900 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/q_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/q_pytest.log | cut -c1-300
