#!/bin/bash
# correctness (new kernel vs the generic kernel) on small/odd shapes + perf on config 2
K=tools/bin/kbench
out=gpurun_out/kb_suite.txt
: > $out
for content in smooth noise dark; do
  for cfg in "96 54 2 1 3 3 1" "131 77 2 1 3 3 2" "37 23 2 1 3 4 3" "64 48 3 2 3 4 2" "129 65 3 2 3 4 1" "96 54 2 1 2 3 2" "200 37 2 1 3 3 1" "960 540 2 1 3 3 2" "1000 300 2 1 3 3 1" "640 360 3 2 3 4 2"; do
    timeout 120 $K $cfg $content 1 0 generic >> $out 2>&1 || echo "   ^^^ rc=$?" >> $out
  done
done
for content in smooth noise; do
  timeout 300 $K 1920 1080 2 1 3 3 32 $content 10 0 generic >> $out 2>&1 || echo "   ^^^ rc=$?" >> $out
  timeout 300 $K 1920 1080 2 1 3 3 32 $content 10 8 >> $out 2>&1 || echo "   ^^^ rc=$?" >> $out
done
timeout 300 $K 2560 1440 3 2 3 4 16 smooth 10 0 generic >> $out 2>&1 || echo "   ^^^ rc=$?" >> $out
timeout 300 $K 3840 2160 2 1 3 3 8 smooth 10 0 generic >> $out 2>&1 || echo "   ^^^ rc=$?" >> $out
cat $out
