#!/usr/bin/env python3
"""Hot-path instruction count of a loop: tools/sass_hot.py OBJ KERNEL_SUBSTR LOOP_START_HEX LOOP_END_HEX
Forward conditional branches that jump over a block containing a CALL (the out-of-line slow paths) are taken:
the skipped block is not counted.  Prints the opcode histogram of what remains."""
import collections, re, subprocess, sys
obj, filt, lo, hi = sys.argv[1], sys.argv[2], int(sys.argv[3], 16), int(sys.argv[4], 16)
txt = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
for f in re.split(r"\n\s*Function : ", txt)[1:]:
    if filt not in f.split("\n", 1)[0]:
        continue
    ins = [(int(a, 16), t) for a, t in re.findall(r"/\*([0-9a-f]{4,6})\*/\s+(.*?);", f)]
    body = [(a, t) for a, t in ins if lo <= a <= hi]
    hist = collections.Counter(); i = 0; n = 0; cold = 0
    while i < len(body):
        a, t = body[i]
        m = re.search(r"^@!?U?P\w+\s+BRA\S*\s+(?:\S+,\s+)?0x([0-9a-f]+)", t)
        if m:
            tgt = int(m.group(1), 16)
            if a < tgt <= hi:
                skipped = [x for x in body if a < x[0] < tgt]
                if any("CALL" in x[1] for x in skipped):
                    hist["BRA"] += 1; n += 1; cold += len(skipped)
                    i += 1 + len(skipped)
                    continue
        op = re.sub(r"^@!?U?P\w+\s+", "", t).split()[0]
        hist[op] += 1; n += 1; i += 1
    print(f"hot {n} instr (cold {cold} skipped): " + " ".join(f"{k}:{v}" for k, v in hist.most_common()))
    break
