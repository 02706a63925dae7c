"""Development aid: FP64 work a kernel executed, by opcode, from the SASS page of an ncu report
(instruction counts are per warp; thread counts exclude predicated-off and inactive lanes).

  python tools/ncu_fp64_counts.py report.ncu-rep
"""
import csv
import io
import subprocess
import sys
from collections import defaultdict

rep = sys.argv[1]
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
kern, hdr, name = [], None, None
for r in rows:
    if r and r[0] == "Kernel Name":
        name = r[1]
    if r and r[0] in ("Address", "#") or (r and "Source" in r and "Instructions Executed" in r):
        hdr = r
        kern.append((name, hdr, []))
        continue
    if hdr is not None and len(r) == len(hdr):
        kern[-1][2].append(dict(zip(hdr, r)))
for name, hdr, body in kern:
    warp, thr = defaultdict(int), defaultdict(int)
    tot = 0
    for d in body:
        op = d["Source"].strip().split()
        if not op:
            continue
        o = op[1] if op[0].startswith("@") else op[0]
        o = o.split(".")[0]
        try:
            n, t = int(d["Instructions Executed"]), int(d["Predicated-On Thread Instructions Executed"])
        except ValueError:
            continue
        tot += n
        if o in ("DFMA", "DMUL", "DADD", "DMMA", "DSETP", "DMNMX"):
            warp[o] += n
            thr[o] += t
    lane_ops = sum(v * (8 if k == "DMMA" else 1) for k, v in thr.items())
    print("kernel %s" % (name or "?")[:90])
    print("  warp instructions %d; FP64-pipe warp instructions %s" % (tot, dict(warp)))
    print("  FP64 lane-operations (DMMA.884 = 8 multiply-adds per lane): %.4e" % lane_ops)
