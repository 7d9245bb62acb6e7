#!/bin/bash
# quick correctness (vs v5) + perf check of the specialised kernels
K=tools/bin/kbench
out=gpurun_out/kb_quick.txt
: > $out
for content in smooth noise dark; do
  for cfg in "96 54 2 1 3 3 1" "64 48 3 2 3 4 2" "96 54 2 1 2 3 2" "960 540 2 1 3 3 2" "640 360 3 2 3 4 2" "1924 270 2 1 3 3 1"; do
    timeout 120 $K $cfg $content 1 0 generic >> $out 2>&1 || echo "   ^^^ rc=$?" >> $out
  done
done
for content in smooth noise; do
  timeout 300 $K 1920 1080 2 1 3 3 32 $content 10 0 generic >> $out 2>&1 || echo "   ^^^ rc=$?" >> $out
  timeout 300 $K 1920 1080 2 1 3 3 32 $content 10 8 >> $out 2>&1 || echo "   ^^^ rc=$?" >> $out
done
timeout 300 $K 2560 1440 3 2 3 4 16 smooth 10 0 generic >> $out 2>&1 || echo "   ^^^ rc=$?" >> $out
timeout 300 $K 2560 1440 3 2 3 4 16 smooth 10 8 >> $out 2>&1 || echo "   ^^^ rc=$?" >> $out
timeout 300 $K 3840 2160 2 1 3 3 8 smooth 10 0 generic >> $out 2>&1 || echo "   ^^^ rc=$?" >> $out
grep -B1 -E "differ|rc=" $out | grep -v "^--" | cut -c1-100,150-330 | grep -E "rc=|differ of|frames=(32|16|8) " 
grep -E "flags=8" $out | cut -c1-260
