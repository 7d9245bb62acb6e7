#!/bin/bash
# same-box A/B of library variants on the headline shape: tools/kb_ab3.sh OUTTAG VARIANT...   ("-" = product library)
K=tools/bin/kbench
out=gpurun_out/kb_ab3_$1.txt; shift
: > $out
for rep in 1 2; do
for v in "$@"; do
  if [ "$v" = "-" ]; then L=lanczos_hls_b200; else L=variants/$v; fi
  for c in smooth noise dark; do
    echo -n "$v rep$rep $c: " >> $out
    LD_LIBRARY_PATH=$L timeout 300 $K 1920 1080 2 1 3 3 64 $c 10 0 2>&1 | grep "frames=64" | sed 's/.*| avg/avg/' | cut -c1-90 >> $out
  done
  echo -n "$v rep$rep tol: " >> $out
  LD_LIBRARY_PATH=$L timeout 300 $K 1920 1080 2 1 3 3 64 smooth 10 8 2>&1 | grep "frames=64" | sed 's/.*| avg/avg/' | cut -c1-90 >> $out
done
done
cat $out
