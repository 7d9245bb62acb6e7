#!/bin/bash
# Turn the ncu reports of tools/final_measure.sh (gpurun_out/prof_<tag>_*.ncu-rep) into the text summaries kept under
# profiles/: key raw metrics, per-opcode executed-instruction table with stall samples, per-region breakdown.
tag=${1:-r02}
for v in exact tol noise dyn; do
  rep=gpurun_out/prof_${tag}_$v.ncu-rep
  [ -f $rep ] || continue
  out=profiles/${tag}_ncu_${v}_full.txt
  python tools/ncu_raw.py $rep > $out
  ncu -i $rep --page source --csv > /tmp/src_$v.csv 2>/dev/null
  echo "" >> $out; echo "--- per-opcode executed warp-instructions (ncu source page)" >> $out
  python tools/ncu_src.py /tmp/src_$v.csv 14 >> $out
  echo "" >> $out; echo "--- code regions (runs of instructions with the same execution count): share of instructions and of stall samples" >> $out
  python tools/ncu_regions.py $rep 12 | awk '{ if ($0 !~ /samples/) print; else { split($0,a,"samples"); split(a[2],b,"%"); if ($0 ~ /total/ || b[1]+0 >= 0.7) print } }' | cut -c1-170 >> $out
done
cp gpurun_out/${tag}_launches.csv profiles/${tag}_ncu_launch_list.csv 2>/dev/null
for f in bench bench_c5 bench_reference_arm; do cp gpurun_out/${tag}_$f.json profiles/${tag}_$f.json 2>/dev/null; done
ls -la profiles | grep ${tag}_
