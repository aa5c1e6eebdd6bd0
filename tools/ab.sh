#!/bin/bash
# A/B session: each argument "tag:ENV=VAL,ENV2=VAL2" runs bench.py (config3, device-timed only) under that environment.
mkdir -p gpurun_out
for spec in "$@"; do
  tag=${spec%%:*}; envs=${spec#*:}
  ( IFS=','; for kv in $envs; do [ -n "$kv" ] && export "$kv"; done
    timeout 300 python bench.py --no-e2e --no-cpu-baseline --steps ${AB_STEPS:-10} ${AB_ARGS} > gpurun_out/ab_${tag}.json 2> gpurun_out/ab_${tag}.err
    python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/ab_${tag}.json").read().strip().splitlines()[-1])
    print("${tag}", "ms/step %.3f" % d["ms_per_step"], "frac %.4f" % d["roofline"]["frac"], "clk", d["clocks"]["sm_mhz"], "bad", d["config"].get("nonzero_status"))
except Exception as e:
    print("${tag}", "FAILED", e, open("gpurun_out/ab_${tag}.err").read()[-500:])
PY
  )
done
