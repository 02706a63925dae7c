"""From an ``ncu --set full --import-source on`` report of the loglr launches of one C3 step: the JSON
bench.py reads (profiles/r2_ncu_counts.json) -- per kernel DRAM bytes, FP64 pipe activity and the FP64
work executed (warp instructions by opcode; a DMMA.884 holds the pipe as long as 8 DFMAs).

  python tools/ncu_profile_json.py report.ncu-rep "<command the report was taken from>" > profiles/r2_ncu_counts.json
"""
import csv
import io
import json
import subprocess
import sys
from collections import OrderedDict, defaultdict

rep, command = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else "")
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], dict(zip(rows[0], rows[1]))
launches = [dict(zip(hdr, r)) for r in rows[2:]]
SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def nbytes(m, key):
    return float(m[key]) * SCALE[units[key]]


src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"],
                     capture_output=True, text=True).stdout
ops, cur = [], None
for r in csv.reader(io.StringIO(src)):
    if r and "Source" in r and "Instructions Executed" in r:
        cur = (r, defaultdict(int))
        ops.append(cur)
        continue
    if cur is None or len(r) != len(cur[0]):
        continue
    d = dict(zip(cur[0], r))
    w = d["Source"].strip().split()
    if not w:
        continue
    o = (w[1] if w[0].startswith("@") else w[0]).split(".")[0]
    try:
        cur[1][o] += int(d["Instructions Executed"])
    except ValueError:
        pass
# (the source page lists every kernel's SASS table once per view: keep one per launch, in order)
ops = ops[::max(1, len(ops) // max(len(launches), 1))][:len(launches)]
out = OrderedDict(command=command, report=rep.split("/")[-1], kernels=[])
for m, (_, o) in zip(launches, ops):
    scalar = sum(o.get(k, 0) for k in ("DFMA", "DMUL", "DADD", "DSETP", "DMNMX"))
    out["kernels"].append(OrderedDict(
        name=m["Kernel Name"][:60], grid=m["launch__grid_size"], duration_us=float(m["gpu__time_duration.sum"]),
        dram_bytes_read=nbytes(m, "dram__bytes_read.sum"), dram_bytes_write=nbytes(m, "dram__bytes_write.sum"),
        sm_cycles=float(m["sm__cycles_elapsed.max"]) if "sm__cycles_elapsed.max" in m else None,
        issue_active_pct=float(m["smsp__issue_active.avg.pct_of_peak_sustained_active"]),
        warp_instructions=int(float(m["smsp__inst_executed.sum"])),
        fp64_scalar_warp_instructions=scalar, dmma_warp_instructions=o.get("DMMA", 0),
        fp64_lane_ops_executed=32.0 * (scalar + 8 * o.get("DMMA", 0)),
        fp64_pipe_busy_frac=((2.0 * scalar + 16.0 * o.get("DMMA", 0)) / (4 * 148 * float(m["sm__cycles_elapsed.max"]))
                             if "sm__cycles_elapsed.max" in m else None)))
print(json.dumps(out, indent=1))
