"""Development aid: per-source-line instruction counts and stall samples from an ncu report.

  python tools/ncu_source_lines.py report.ncu-rep [kernel_index] [top]
"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
kidx = int(sys.argv[2]) if len(sys.argv) > 2 else 0
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
# sections: "File Path" rows; a kernel's sections follow each other; kernels repeat the first file
secs, cur = [], None
for r in rows:
    if r and r[0] == "File Path":
        cur = {"file": r[1], "rows": []}
        secs.append(cur)
    elif cur is not None:
        cur["rows"].append(r)
first = secs[0]["file"]
kern, k = [], -1
for s in secs:
    if s["file"] == first:
        k += 1
        kern.append([])
    kern[k].append(s)
def num(x):
    try:
        return int(x)
    except ValueError:
        return 0


tot_inst = tot_samp = 0
lines = {}
for s in kern[kidx]:
    hdr = None
    line_no, line_src = None, None
    for r in s["rows"]:
        if r and r[0] == "Line No":
            hdr = r
            continue
        if hdr is None or len(r) < len(hdr):
            continue
        d = dict(zip(hdr, r))
        # (two "Source" columns: the zip keeps the SASS one)
        if r[0] != "":
            line_no, line_src = r[0], r[1]
            continue
        key = (s["file"].split("/")[-1], int(line_no))
        e = lines.setdefault(key, {"src": line_src, "inst": 0, "samp": 0, "fp64": 0, "stalls": {}})
        n = num(d["Instructions Executed"])
        e["inst"] += n
        e["samp"] += num(d["# Samples"])
        if r[3].strip().split()[0] in ("DFMA", "DMUL", "DADD", "DSETP", "DMNMX") or r[3].strip().startswith("@") and " D" in r[3][:14]:
            e["fp64"] += n
        for kk, v in d.items():
            if kk.startswith("stall_") and "Not Issued" not in kk and v not in ("", "0"):
                e["stalls"][kk[6:]] = e["stalls"].get(kk[6:], 0) + num(v)
        tot_inst += n
        tot_samp += num(d["# Samples"])
print("kernel %d: %d warp instructions, %d samples" % (kidx, tot_inst, tot_samp))
for key, e in sorted(lines.items(), key=lambda kv: -kv[1]["samp"])[:top]:
    st = sorted(e["stalls"].items(), key=lambda kv: -kv[1])[:3]
    print("%-14s %5d  samp %5.1f%%  inst %5.1f%%  fp64 %4.0f%%  %-38s %s" % (
        key[0][:14], key[1], 100.0 * e["samp"] / max(tot_samp, 1), 100.0 * e["inst"] / max(tot_inst, 1),
        100.0 * e["fp64"] / max(e["inst"], 1), " ".join("%s:%d" % x for x in st), e["src"].strip()[:90]))
