#!/usr/bin/env python3
"""Per-region summary of an ncu report's source page: tools/ncu_regions.py REPORT.ncu-rep [min_instr]
Instructions are grouped into runs with (nearly) the same execution count -- loop bodies, slow paths, glue --
and each run is listed with its share of executed warp-instructions and of warp-stall samples (~ time)."""
import csv, io, subprocess, sys
rep = sys.argv[1]
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
hdr = rows[1]
ia, isrc, isamp, iex, ith = (hdr.index(k) for k in ("Address", "Source", "# Samples", "Instructions Executed", "Thread Instructions Executed"))
data = []
base = None
for r in rows[2:]:
    try:
        a = int(r[ia], 16)
        base = a if base is None else base
        data.append((a - base, r[isrc].strip(), int(r[isamp]), int(r[iex]), int(r[ith])))
    except (ValueError, IndexError):
        pass
tot = sum(d[3] for d in data); tots = sum(d[2] for d in data)
print(rows[0][1][:120]); print(f"total warp-inst {tot}  samples {tots}")
i = 0
while i < len(data):
    j = i
    while j + 1 < len(data) and abs(data[j + 1][3] - data[i][3]) <= max(2000, 0.03 * data[i][3]): j += 1
    n = j - i + 1
    ex = sum(d[3] for d in data[i:j + 1]); sm = sum(d[2] for d in data[i:j + 1]); th = sum(d[4] for d in data[i:j + 1])
    if ex and (n >= int(sys.argv[2]) if len(sys.argv) > 2 else n >= 6):
        print(f"{data[i][0]:#07x}-{data[j][0]:#07x} n={n:4d} ex/instr {data[i][3]:9d} inst {100*ex/tot:5.1f}%  samples {100*sm/tots:5.1f}%  thr {th/max(ex,1):4.1f}   {data[i][1][:40]}")
    i = j + 1
