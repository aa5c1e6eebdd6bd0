"""gpurun_out/p_* (tools/profile_session.sh) -> the committed summaries under profiles/ (round tag r02)."""
import csv, io, json, os, shutil, subprocess, sys

TAG = sys.argv[1] if len(sys.argv) > 1 else "r02"
KEYS = ["Kernel Name", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_warps", "gpu__time_duration.sum",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "sass__inst_executed_local_loads", "sass__inst_executed_local_stores", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "sm__icc_request_hit_rate.pct",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio"]
SCALE = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
TIME = {"ms": 1.0, "us": 1e-3, "ns": 1e-6, "s": 1e3}


def raw(path):
    rows = list(csv.reader(open(path)))
    return {h: (u, v) for h, u, v in zip(rows[0], rows[1], rows[2])}


def summarise(name, kernel, what, units=None, bytes_per_unit=None, unit_name="pose"):
    rp = f"gpurun_out/p_{name}_raw.csv"
    if not os.path.exists(rp) or sum(1 for _ in open(rp)) < 3:
        print("no capture for", name)
        return None
    m = raw(rp)
    num = lambda k: float(m[k][1].replace(",", ""))
    out = f"profiles/{TAG}_{name}_kernel_metrics.txt"
    with open(out, "w") as f:
        f.write(f"# ncu --set full --clock-control none --import-source on, one launch of {kernel}: {what}\n")
        f.write("# (cold-cache, serialised replay: use shares and byte counts, not the absolute duration)\n")
        for k in KEYS:
            if k in m:
                f.write(f"{k:82s} {m[k][1]} {m[k][0]}\n")
        rd = num("dram__bytes_read.sum") * SCALE[m["dram__bytes_read.sum"][0]]
        wr = num("dram__bytes_write.sum") * SCALE[m["dram__bytes_write.sum"][0]]
        dur = num("gpu__time_duration.sum") * TIME[m["gpu__time_duration.sum"][0]]
        f.write(f"\ndram read + write                          {(rd + wr) / 1e9:.3f} GB in {dur:.3f} ms = {(rd + wr) / dur / 1e6:.0f} GB/s\n")
        if units:
            f.write(f"per {unit_name}: {(rd + wr) / units:.1f} B of DRAM traffic, {num('smsp__inst_executed.sum') * num('smsp__thread_inst_executed_per_inst_executed.ratio') / units:.0f} thread-instructions\n")
            if bytes_per_unit:
                f.write(f"algorithmic bytes ({bytes_per_unit} B/{unit_name})           {units * bytes_per_unit / 1e9:.3f} GB -> traffic = {100 * (rd + wr) / (units * bytes_per_unit):.1f} % of algorithmic, {units * bytes_per_unit / dur / 1e6:.0f} GB/s algorithmic under ncu\n")
    sp = f"gpurun_out/p_{name}_source.csv"
    if os.path.exists(sp):
        r = subprocess.run([sys.executable, "tools/ncu_lines.py", sp, "gps_optimize_slam_b200/libgsf.so", kernel, "40"], capture_output=True, text=True)
        open(f"profiles/{TAG}_{name}_kernel_hot_lines.txt", "w").write(r.stdout)
    return {"dram_bytes": rd + wr, "ms": dur}


fast = summarise("fast", "fuse_fast_kernel", "65 536 trajectories x 1000 poses (config 3 slab)", 65536 * 1000, 144)
summarise("general", "fuse_traj_kernel", "the deferred half of 65 536 x 1000 poses with 50 % outage trajectories (general kernel + RTS)", 32768 * 1000, 144)
summarise("combine", "grid_combine_kernel", "config 5, 262 144 hypotheses x 4541 poses (4493 evaluated)", 262144 * 4493, None, "query")
summarise("ate", "ate_nn_kernel", "131 072 trajectories x 1000 poses", 131072 * 1000, 56)
summarise("f32", "fuse_f32_kernel", "fp32 mode, 131 072 trajectories x 1000 poses", 131072 * 1000, 72)
summarise("assoc_m", "assoc_long_moments_kernel", "config 4 at 2e7 samples: local-halo spline moments", 20000000, 56, "sample")
summarise("assoc_e", "assoc_long_eval_kernel", "config 4 at 2e7 samples: per-stamp evaluation", 20000000, 89, "sample")
summarise("warp", "fuse_warp_kernel", "65 536 x 271 poses, one warp per trajectory (GSF_FAST_CT=1; measured alternative, not the default)", 65536 * 271, 144)
for c in (2, 3, 4, 5):
    src = f"gpurun_out/p_bench_config{c}.json"
    if os.path.exists(src) and os.path.getsize(src) > 0:
        shutil.copy(src, f"profiles/{TAG}_bench_config{c}.json")
if os.path.exists("gpurun_out/p_launches_config3.csv"):
    shutil.copy("gpurun_out/p_launches_config3.csv", f"profiles/{TAG}_launches_config3.csv")
# measured DRAM bytes of the full-size timed launch
p = "gpurun_out/p_dram_fullsize.csv"
if os.path.exists(p):
    rows = [r for r in csv.reader(open(p)) if r and not r[0].startswith("==")]
    hdr = rows[0]; vals = {}
    for r in rows[1:]:
        d = dict(zip(hdr, r))
        vals[d["Metric Name"]] = (float(d["Metric Value"].replace(",", "")), d["Metric Unit"])
    shutil.copy(p, f"profiles/{TAG}_dram_fullsize_launch.csv")
    rd = vals["dram__bytes_read.sum"][0] * SCALE[vals["dram__bytes_read.sum"][1]]
    wr = vals["dram__bytes_write.sum"][0] * SCALE[vals["dram__bytes_write.sum"][1]]
    B, n = 1 << 20, 1000
    json.dump({"workload": "config3", "kernel": "fuse_fast_kernel", "profiled_trajectories": B, "poses_per_trajectory": n,
               "dram_bytes_per_trajectory": (rd + wr) / B, "dram_bytes_profiled_launch": rd + wr,
               "source": f"profiles/{TAG}_dram_fullsize_launch.csv (ncu, dram__bytes_read.sum + dram__bytes_write.sum of the timed launch shape: 2^20 trajectories x 1000 poses, one pass, no replay)"},
              open("profiles/traffic.json", "w"), indent=1)
    print("full-size launch: %.2f GB DRAM traffic = %.1f B/pose (%.1f %% of the 144 B algorithmic)" % ((rd + wr) / 1e9, (rd + wr) / (B * n), 100 * (rd + wr) / (B * n * 144)))
print(os.listdir("profiles"))
