#!/bin/bash
# A/B of library builds: ./tools/ab_libs.sh v1 v2 ...  (gps_optimize_slam_b200/libgsf_<tag>.so; "base" = libgsf.so)
mkdir -p gpurun_out
P=$PWD/gps_optimize_slam_b200
for tag in "$@"; do
  lib=$P/libgsf_${tag}.so; [ "$tag" = base ] && lib=$P/libgsf.so
  export GSF_LIB=$lib
  if [ -z "$AB_NOTEST" ]; then
    timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "fused or fast_kernel or reproducible or full_size or device_generator" > gpurun_out/ab_${tag}_pytest.log 2>&1
    echo "$tag pytest: $(tail -1 gpurun_out/ab_${tag}_pytest.log)"
  fi
  timeout 300 python bench.py --no-e2e --no-cpu-baseline --steps ${AB_STEPS:-10} ${AB_ARGS} > gpurun_out/ab_${tag}.json 2> gpurun_out/ab_${tag}.err
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/ab_${tag}.json").read().strip().splitlines()[-1])
    print("${tag}", "ms/step %.3f" % d["ms_per_step"], "frac %.4f" % d["roofline"]["frac"], "clk", d["clocks"]["sm_mhz"], "bad", d["config"].get("nonzero_status"))
except Exception as e:
    print("${tag}", "FAILED", e, open("gpurun_out/ab_${tag}.err").read()[-500:])
PY
done
