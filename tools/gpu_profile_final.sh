#!/bin/bash
# Round profile artefacts: launch lists of the bench commands + full capture of the fast fused kernel.
mkdir -p gpurun_out
BENCH="python bench.py --steps 3 --warmup 3 --no-cpu-baseline"
$BENCH > gpurun_out/pf_bench_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/pf_launches.csv $BENCH > gpurun_out/pf_ncu_launch.log 2>&1
tail -1 gpurun_out/pf_bench_plain.log | cut -c1-300
CMD="python bench.py --steps 2 --warmup 3 --trajectories 65536 --no-cpu-baseline --no-e2e"
$CMD > gpurun_out/pf_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:fuse_fast -s 3 -c 1 -f -o gpurun_out/pf_fast $CMD > gpurun_out/pf_ncu_full.log 2>&1
tail -2 gpurun_out/pf_ncu_full.log
for wl in config4 config5; do
C="python bench.py --workload $wl --steps 2 --warmup 3 --no-cpu-baseline"
$C > gpurun_out/pf_${wl}_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active --clock-control none -c 120 --csv --log-file gpurun_out/pf_${wl}_launches.csv $C > gpurun_out/pf_${wl}_ncu.log 2>&1
tail -1 gpurun_out/pf_${wl}_plain.log | cut -c1-200
done
