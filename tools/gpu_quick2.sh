#!/bin/bash
for t in 128; do
echo "threads=$t"
GSF_FUSE_THREADS=$t timeout 600 python bench.py --steps 5 --warmup 3 --trajectories 65536 --no-cpu-baseline --no-e2e 2>&1 | tail -1 | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: print(l[:300]); continue
    print('value %.3e'%d['value'], 'ms/step %.3f'%d['ms_per_step'], 'frac %.3f'%d['roofline']['frac'], 'bad', d['config']['nonzero_status'])
"
done
