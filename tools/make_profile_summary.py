"""Turn gpurun_out/<tag>.ncu-rep into the committed summaries under profiles/.

    python tools/make_profile_summary.py gpurun_out/p2_fuse.ncu-rep r01 65536 1000
"""
import csv
import io
import json
import subprocess
import sys

rep, tag, trajs, n = sys.argv[1], sys.argv[2], int(sys.argv[3]), int(sys.argv[4])
kname = sys.argv[5] if len(sys.argv) > 5 else "fuse_fast_kernel"
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
m = {h: (u, v) for h, u, v in zip(hdr, units, vals)}
keys = ["Kernel Name", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_warps",
        "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "sass__inst_executed_local_loads", "sass__inst_executed_local_stores",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "sm__icc_request_hit_rate.pct", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio"]
def num(k):
    return float(m[k][1].replace(",", ""))
scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
rd = num("dram__bytes_read.sum") * scale[m["dram__bytes_read.sum"][0]]
wr = num("dram__bytes_write.sum") * scale[m["dram__bytes_write.sum"][0]]
dur_ms = num("gpu__time_duration.sum") * {"ms": 1.0, "us": 1e-3, "ns": 1e-6, "s": 1e3}[m["gpu__time_duration.sum"][0]]
with open(f"profiles/{tag}_fuse_metrics.txt", "w") as f:
    f.write(f"# ncu --set full --clock-control none, one launch of {kname} over {trajs} trajectories x {n} poses\n")
    f.write("# (cold-cache, serialised replay: use shares and byte counts, not the absolute duration)\n")
    for k in keys:
        if k in m:
            f.write(f"{k:70s} {m[k][1]} {m[k][0]}\n")
    alg = trajs * n * 144
    f.write(f"\nalgorithmic bytes (144 B/pose)            {alg / 1e9:.3f} GB\n")
    f.write(f"dram read + write                          {(rd + wr) / 1e9:.3f} GB  ({(rd + wr) / (trajs * n):.1f} B/pose, {100 * (rd + wr) / alg:.1f} % of algorithmic)\n")
    f.write(f"achieved under ncu                         {alg / (dur_ms * 1e-3) / 1e9:.0f} GB/s algorithmic\n")
json.dump({"workload": "config3", "kernel": kname, "profiled_trajectories": trajs, "poses_per_trajectory": n,
           "dram_bytes_per_trajectory": (rd + wr) / trajs, "dram_bytes_profiled_launch": rd + wr,
           "source": f"profiles/{tag}_fuse_metrics.txt (ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum)"},
          open("profiles/traffic.json", "w"), indent=1)
print(open(f"profiles/{tag}_fuse_metrics.txt").read())
