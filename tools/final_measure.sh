#!/bin/bash
# round-end evidence: bench lines of BASELINE configs 2-5 + the reference arm, ncu launch list and one full capture
# (each ncu pass only after the same command exited 0 without ncu).  Outputs under gpurun_out/<tag>_*.
tag=${1:-r01d}
o=gpurun_out
python bench.py > $o/${tag}_bench_c2.json 2> $o/${tag}_err.txt || exit 1
python bench.py --workload c3 --no-cpu-baseline > $o/${tag}_bench_c3.json 2>> $o/${tag}_err.txt
python bench.py --workload c4 --no-cpu-baseline > $o/${tag}_bench_c4.json 2>> $o/${tag}_err.txt
python bench.py --workload c5 --steps 5 --warmup 3 --no-cpu-baseline > $o/${tag}_bench_c5.json 2>> $o/${tag}_err.txt
python bench.py --impl reference --steps 4 --warmup 1 > $o/${tag}_bench_reference_arm.json 2>> $o/${tag}_err.txt
python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > $o/${tag}_plain.log 2>&1 || exit 2
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $o/${tag}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > $o/${tag}_ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:lanczos_v6 -s 4 -c 1 -f -o $o/prof_${tag}_exact \
    python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > $o/${tag}_ncu_full.log 2>&1
tools/bin/kbench 1920 1080 2 1 3 3 32 smooth 1 8 > $o/${tag}_plain_tol.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:lanczos_v6 -s 3 -c 1 -f -o $o/prof_${tag}_tol \
    tools/bin/kbench 1920 1080 2 1 3 3 32 smooth 1 8 > $o/${tag}_ncu_tol.log 2>&1
for f in c2 c3 c4 c5; do python - <<PY
import json
d=json.loads(open("$o/${tag}_bench_$f.json").read().strip().splitlines()[-1])
print("$f", round(d["value"]), "Mpix/s frac", round(d["roofline"]["frac"],3), "e2e", round(d["e2e"]["value"]) if d.get("e2e") else None,
      "tol", round((d.get("tolerance_mode") or {}).get("value") or 0), "worst", round((d.get("worst_case") or {}).get("value") or 0), d["clocks"])
PY
done
cut -c1-250 $o/${tag}_bench_reference_arm.json
