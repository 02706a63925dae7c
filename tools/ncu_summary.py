"""Summarise an .ncu-rep (raw + source pages) into a short text report for profiles/."""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
for vals in rows[2:]:
    m = dict(zip(hdr, zip(units, vals)))
    print("kernel:", m["Kernel Name"][1], " grid", m["launch__grid_size"][1], " block", m["launch__block_size"][1])
    keys = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__occupancy_limit_registers",
            "launch__occupancy_limit_shared_mem", "launch__waves_per_multiprocessor",
            "sm__warps_active.avg.pct_of_peak_sustained_active",
            "sm__cycles_elapsed.avg.per_second", "smsp__inst_executed.sum",
            "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
            "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
            "smsp__issue_active.avg.pct_of_peak_sustained_active",
            "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
            "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
            "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
            "sm__throughput.avg.pct_of_peak_sustained_elapsed",
            "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
            "lts__t_bytes.sum", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
            "smsp__average_warp_latency_per_inst_issued.ratio"]
    for k in keys:
        if k in m:
            print("  %-68s %s %s" % (k, m[k][1], m[k][0]))
    st = []
    for k, (u, v) in m.items():
        if k.startswith("smsp__average_warps_issue_stalled_") and k.endswith("_per_issue_active.ratio"):
            try:
                st.append((float(v), k[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]))
            except ValueError:
                pass
    print("  stall reasons (warps stalled per issue-active cycle):")
    for v, k in sorted(st, reverse=True)[:8]:
        print("    %-28s %.3f" % (k, v))
