#!/usr/bin/env python3
"""Opcode histogram of the innermost backward-branch loops of every kernel in a cubin/executable
(cuobjdump -sass): tools/sass_loops.py BINARY [name-substring]. Development aid for the ubench files
and for counting instructions per loop body of the product kernels."""
import collections, re, subprocess, sys

def main():
    binary = sys.argv[1]
    filt = sys.argv[2] if len(sys.argv) > 2 else ""
    txt = subprocess.run(["cuobjdump", "-sass", binary], capture_output=True, text=True).stdout
    funcs = re.split(r"\n\s*Function : ", txt)[1:]
    for f in funcs:
        name = f.split("\n", 1)[0].strip()
        if filt and filt not in name:
            continue
        ins = re.findall(r"/\*([0-9a-f]{4,6})\*/\s+(.*?);", f)
        addr = [int(a, 16) for a, _ in ins]
        loops = []
        for a, t in ins:
            m = re.search(r"\bBRA\S*\s+(?:\S+,\s+)?0x([0-9a-f]+)", t)
            if m and int(m.group(1), 16) <= int(a, 16):
                loops.append((int(m.group(1), 16), int(a, 16)))
        print("==", subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip()[:150], f"({len(ins)} instr)")
        for lo, hi in loops:
            body = [t for a, t in zip(addr, (t for _, t in ins)) if lo <= a <= hi]
            if len(body) < 8:
                continue
            hist = collections.Counter()
            for t in body:
                t = re.sub(r"^@!?U?P\w+\s+", "", t)
                hist[t.split()[0]] += 1
            print(f"   loop {lo:#x}..{hi:#x} {len(body)} instr: " + " ".join(f"{k}:{v}" for k, v in hist.most_common()))

main()
