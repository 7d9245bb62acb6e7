#!/bin/bash
# round-2 A/B: correctness vs the generic kernel + perf of the headline shapes; tools/kb_r02.sh TAG [LIBDIR]
tag=$1; L=${2:-lanczos_hls_b200}
K=tools/bin/kbench
out=gpurun_out/r02_kb_$tag.txt
: > $out
export LD_LIBRARY_PATH=$L
for c in smooth noise dark; do timeout 300 $K 1920 1080 2 1 3 3 64 $c 10 0 generic >> $out 2>&1 || echo "   ^^^ rc=$?" >> $out; done
timeout 300 $K 1920 1080 2 1 3 3 64 smooth 10 8 generic >> $out 2>&1
timeout 300 $K 2560 1440 3 2 3 4 32 smooth 10 0 generic >> $out 2>&1 || echo "   ^^^ rc=$?" >> $out
timeout 300 $K 2560 1440 3 2 3 4 32 noise 5 0 generic >> $out 2>&1 || echo "   ^^^ rc=$?" >> $out
timeout 300 $K 3840 2160 2 1 3 3 16 smooth 10 0 >> $out 2>&1
for cfg in "96 54 2 1 3 3 1" "131 77 2 1 3 3 2" "37 23 2 1 3 4 3" "129 65 3 2 3 4 1" "96 54 2 1 2 3 2" "1000 300 2 1 3 3 1" "640 360 3 2 3 4 2" "1924 270 2 1 3 3 1"; do
  for c in noise dark; do timeout 120 $K $cfg $c 1 0 generic >> $out 2>&1 || echo "   ^^^ rc=$?" >> $out; done
done
grep -v "back to back\|streams" $out | cut -c1-120,150-330
