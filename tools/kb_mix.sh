#!/bin/bash
# same-box A/B of library variants on the headline shape incl. the 50/50 mix: tools/kb_mix.sh OUTTAG VARIANT...  ("-" = product)
K=tools/bin/kbench
out=gpurun_out/kb_mix_$1.txt; shift
: > $out
for rep in 1 2; do
for v in "$@"; do
  if [ "$v" = "-" ]; then L=lanczos_hls_b200; else L=variants/$v; fi
  for c in smooth noise mix; do
    echo -n "$v rep$rep $c: " >> $out
    LD_LIBRARY_PATH=$L timeout 300 $K 1920 1080 2 1 3 3 64 $c 10 0 2>&1 | grep "frames=64" | sed 's/.*| avg/avg/' | cut -c1-90 >> $out
  done
  echo -n "$v rep$rep 4K smooth: " >> $out
  LD_LIBRARY_PATH=$L timeout 300 $K 3840 2160 2 1 3 3 16 smooth 10 0 2>&1 | grep "frames=16" | sed 's/.*| avg/avg/' | cut -c1-90 >> $out
done
done
cat $out
