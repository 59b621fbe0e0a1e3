"""Per-instruction budget of a kernel from the source page of an ncu report (--set full
--import-source on): opcode histogram, instructions per work unit, the mbarrier wait loops with their
retry counts and stall samples, and the stall reasons.  This is where the warp-skew finding of DESIGN.md
section 4.1 came from.  Usage: python tools/ncu_budget.py r3n knn2 [units]   (units = work units of the
captured launch, default: warp-tiles of the default bench workload) -> profiles/<tag>_<kernel>_budget.txt"""
import collections
import csv
import io
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag, kern = sys.argv[1], sys.argv[2]
# default bench launch: 19,900 pairs x 32 query blocks x 64 train tiles x 16 epilogue warps
units = float(sys.argv[3]) if len(sys.argv) > 3 else 19900 * 32 * 64 * 16
rep = os.path.join(ROOT, "gpurun_out", f"{tag}_{kern}.ncu-rep")
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
h = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
# a report may hold several launches (one source page each): take the first one
data = []
for r in rows[h + 1:]:
    if r and r[0] == "Address":
        break
    if len(r) == len(rows[h]):
        data.append(r)
hdr = rows[h]
col = {name: hdr.index(name) for name in hdr}
ex = [int(r[col["Instructions Executed"]]) for r in data]
smp = [int(r[col["# Samples"]]) for r in data]
src = [r[col["Source"]].strip() for r in data]


def opcode(s):
    t = s.split()
    return (t[1] if t[0].startswith("@") else t[0]).split(".")[0]


out = [f"# {rep.replace(ROOT + '/', '')}: {len(data)} SASS lines, {sum(ex):.4g} warp instructions, "
       f"{sum(smp)} stall samples, {units:.4g} work units (epilogue warp-tiles)",
       f"instructions per unit (all warps of the CTA): {sum(ex) / units:.1f}", "", "opcode  per-unit  share"]
hist = collections.Counter()
for s, e in zip(src, ex):
    hist[opcode(s)] += e
for op, v in hist.most_common(22):
    out.append(f"{op:12s} {v / units:8.2f} {100 * v / sum(ex):6.2f}%")
out += ["", "mbarrier waits (line, executions per unit, stall samples, instruction):"]
for i, (s, e, m) in enumerate(zip(src, ex, smp)):
    if ("TRYWAIT" in s or "NANOSLEEP" in s) and (e > 1e-3 * units or m > 1e-3 * sum(smp)):
        out.append(f"{i:5d} {e / units:8.3f} {m:8d}  {s[:70]}")
out += ["", "stall reasons (share of all samples):"]
for name in hdr:
    if name.startswith("stall_") and "Not Issued" not in name:
        v = sum(int(r[col[name]]) for r in data)
        if v > 0.005 * sum(smp):
            out.append(f"{name[6:]:20s} {100 * v / sum(smp):5.1f}%")
path = os.path.join(ROOT, "profiles", f"{tag}_{kern}_budget.txt")
open(path, "w").write("\n".join(out) + "\n")
print("\n".join(out))
