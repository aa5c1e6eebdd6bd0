#!/bin/bash
# Round-end profiling session: bench lines of every workload, ncu launch list, ncu --set full captures of the main kernels.
mkdir -p gpurun_out
B="--no-cpu-baseline"
timeout 300 python bench.py --workload config2 $B --no-mixed --no-config5 > gpurun_out/p_bench_config2.json 2> gpurun_out/p_c2.err
timeout 600 python bench.py --workload config4 > gpurun_out/p_bench_config4.json 2> gpurun_out/p_c4.err
timeout 600 python bench.py --workload config5 > gpurun_out/p_bench_config5.json 2> gpurun_out/p_c5.err
timeout 900 python bench.py > gpurun_out/p_bench_config3.json 2> gpurun_out/p_c3.err
echo "benches done: $(ls gpurun_out/p_bench_*.json | wc -l)"
# launch list of the default command (shares of the step, cold-cache serialised times)
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/p_launches_config3.csv python bench.py $B --no-e2e --steps 2 --warmup 3 > /dev/null 2>&1
# measured DRAM bytes of the full-size timed launch (one pass: no replay, no save/restore of the 151 GB)
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:fuse_fast_kernel -s 3 -c 1 --csv --log-file gpurun_out/p_dram_fullsize.csv python bench.py $B --no-e2e --no-mixed --no-config5 --ate-trajectories 1024 --steps 2 --warmup 3 > /dev/null 2>&1
./tools/ncu_kernel.sh p_fast fuse_fast_kernel 2 -- python tools/ab_short.py 65536 1000 3
./tools/ncu_kernel.sh p_general fuse_traj_kernel 3 -- python tools/fast_vs_general.py 65536 1000 0.5
./tools/ncu_kernel.sh p_combine grid_combine_kernel 3 -- python bench.py --workload config5 $B --steps 1 --warmup 3
./tools/ncu_kernel.sh p_ate ate_nn_kernel 2 -- python bench.py $B --no-e2e --no-config5 --no-mixed --steps 1 --trajectories 131072
GSF_FAST_CT=1 ./tools/ncu_kernel.sh p_warp fuse_warp_kernel 2 -- python tools/ab_short.py 65536 271 3
./tools/ncu_kernel.sh p_f32 fuse_f32_kernel 2 -- python tools/f32_bench.py 131072 1000
./tools/ncu_kernel.sh p_assoc_m assoc_long_moments_kernel 2 -- python bench.py --workload config4 $B --steps 1 --warmup 3 --poses 20000000
./tools/ncu_kernel.sh p_assoc_e assoc_long_eval_kernel 2 -- python bench.py --workload config4 $B --steps 1 --warmup 3 --poses 20000000
rm -f gpurun_out/p_warp.ncu-rep gpurun_out/p_assoc_m.ncu-rep gpurun_out/p_assoc_e.ncu-rep gpurun_out/p_f32.ncu-rep gpurun_out/p_ate.ncu-rep
ls -la gpurun_out | tail -30
