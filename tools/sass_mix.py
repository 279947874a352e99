"""Summarise an `ncu --page source --csv` export: executed warp-instructions per SASS opcode."""
import collections
import csv
import re
import sys

path, frames = sys.argv[1], float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
rows = list(csv.reader(open(path)))
hdr = next(r for r in rows if "Source" in r and "Instructions Executed" in r)
idx = {h: i for i, h in enumerate(hdr)}
ops, samples, tot = collections.Counter(), collections.Counter(), 0
for r in rows:
    if len(r) < len(hdr) or r is hdr or not r[idx["Instructions Executed"]].isdigit():
        continue
    sass = r[idx["Source"]].strip()
    n = int(r[idx["Instructions Executed"]])
    m = re.match(r"(@!?U?P\d+\s+)?([A-Z0-9_.]+)", sass)
    op = m.group(2).split(".")[0] if m else "?"
    ops[op] += n
    tot += n
    samples[op] += int(r[idx["# Samples"]])
print("total warp-instructions", tot, " per unit", round(tot / frames, 1))
for op, n in ops.most_common(30):
    print(f"{op:10s} {n:12d} {100 * n / tot:5.1f}%  per-unit {n / frames:8.1f}   stall-samples {samples[op]}")
