#!/bin/bash
# round evidence on ONE B200 (run under gpurun): the default bench line (which carries c3 / c4 / c5_bands), the
# reference arm, the ncu launch list of the same command and full captures of the headline kernel on image-like
# content (exact and tolerance) and on uniform noise.  Every ncu pass runs only after the same command exited 0
# without ncu.  Outputs under gpurun_out/<tag>_*; summarise them here with tools/summarise_profiles.sh <tag>.
tag=${1:-r02}
o=gpurun_out
python bench.py > $o/${tag}_bench.json 2> $o/${tag}_err.txt || exit 1
python bench.py --workload c5 --steps 5 --warmup 3 > $o/${tag}_bench_c5.json 2>> $o/${tag}_err.txt
python bench.py --impl reference --steps 4 --warmup 1 > $o/${tag}_bench_reference_arm.json 2>> $o/${tag}_err.txt
python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > $o/${tag}_plain.log 2>&1 || exit 2
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $o/${tag}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > $o/${tag}_ncu_launches.log 2>&1
K=tools/bin/kbench
for v in "exact smooth 0" "tol smooth 8" "noise noise 0"; do
  set -- $v
  $K 1920 1080 2 1 3 3 64 $2 1 $3 > $o/${tag}_plain_$1.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:lanczos_v6 -s 3 -c 1 -f -o $o/prof_${tag}_$1 \
      $K 1920 1080 2 1 3 3 64 $2 1 $3 > $o/${tag}_ncu_$1.log 2>&1
done
$K 16384 4096 17 10 3 3 1 smooth 1 0 > $o/${tag}_plain_dyn.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:lanczos_dyn -s 3 -c 1 -f -o $o/prof_${tag}_dyn \
    $K 16384 4096 17 10 3 3 1 smooth 1 0 > $o/${tag}_ncu_dyn.log 2>&1
python - <<PY
import json
d=json.loads(open("$o/${tag}_bench.json").read().strip().splitlines()[-1])
print("value", round(d["value"]), "frac", round(d["roofline"]["frac"],3), {k:(round(v["value"]), round(v["roofline_frac"],3)) for k,v in d["by_content"].items()})
print("tol", round(d["tolerance_mode"]["value"]), round(d["tolerance_mode"]["roofline_frac"],3), d["tolerance_mode"]["exact_match_fraction"])
print("e2e", round(d["e2e"]["value"]), "of ceiling", round(d["e2e"]["frac_of_pcie_ceiling"],3), "cpu", d["cpu_baseline"] and round(d["cpu_baseline"]["value"],1), d["clocks"])
for k in ("c3","c4","c5_bands"): print(k, {c:(round(v["value"]), round(v.get("roofline_frac", v.get("roofline_frac_per_gpu")),3)) for c,v in d[k]["by_content"].items()})
PY
cut -c1-250 $o/${tag}_bench_reference_arm.json
