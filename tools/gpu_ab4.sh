#!/bin/bash
mkdir -p gpurun_out
{
timeout 200 python tools/fast_vs_general.py 65536 1000 2>&1 | tail -3
timeout 200 python tools/fast_vs_general.py 65536 271 | tail -3
timeout 200 python tools/fast_vs_general.py 2048 1000 0.02 | tail -3
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
} > gpurun_out/ab.log 2>&1
cat gpurun_out/ab.log
