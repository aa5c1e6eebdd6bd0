#!/bin/bash
# quick perf + parity loop
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/q_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/q_pytest.log | cut -c1-200
for wl in "--trajectories 65536" "--workload config2"; do
timeout 600 python bench.py --steps 5 --warmup 3 $wl --no-cpu-baseline --no-e2e 2>&1 | tail -1 | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: print(l[:300]); continue
    print(d['config']['workload'][:40], 'value %.3e'%d['value'], 'ms/step %.3f'%d['ms_per_step'], 'frac %.3f'%d['roofline']['frac'], 'GB/s %.0f'%d['roofline']['achieved'], d['clocks'])
"
done
