#!/usr/bin/env python3
"""Summarise an `ncu --page source --csv` export: opcode histogram (executed warp-instructions),
stall-sample totals, and the hottest instructions.  usage: ncu_src.py src.csv [top_n]"""
import csv, sys, collections, re
path = sys.argv[1]; topn = int(sys.argv[2]) if len(sys.argv) > 2 else 25
rows = list(csv.reader(open(path)))
hdr = rows[1]
ia, isrc, isamp, iex = hdr.index("Address"), hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
ops = collections.Counter(); samp = collections.Counter(); stalls = collections.Counter()
tot = 0; tots = 0
body = []
for r in rows[2:]:
    if len(r) <= iex: continue
    try: ex = int(r[iex]); sm = int(r[isamp])
    except ValueError: continue
    src = r[isrc]
    m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)", src)
    op = m.group(2) if m else src
    base = op.split(".")[0]
    if base in ("IMAD", "LDS", "STS", "STG"):
        base = ".".join(op.split(".")[:2])
    ops[base] += ex; samp[base] += sm; tot += ex; tots += sm
    for i in stall_cols:
        try: stalls[hdr[i]] += int(r[i])
        except ValueError: pass
    body.append((sm, ex, r[ia], src))
print(f"total warp-inst {tot}  samples {tots}")
for op, n in ops.most_common(45):
    print(f"{op:22s} {n:12d} {100*n/tot:6.2f}%   samples {100*samp[op]/max(tots,1):5.1f}%")
print({k: v for k, v in stalls.most_common(10)})
print("hottest by samples:")
for sm, ex, a, src in sorted(body, reverse=True)[:topn]:
    print(f"{sm:6d} {ex:9d} {a[-5:]} {src}")
