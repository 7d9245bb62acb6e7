#!/bin/bash
# dynamic-phase kernels (lanczos_dyn.cu): correctness against the generic kernel on small shapes + timing of the 17/10 shapes
#   tools/kb_dyn.sh TAG [LIBDIR]
tag=$1; L=${2:-lanczos_hls_b200}
K=tools/bin/kbench
out=gpurun_out/kb_dyn_$tag.txt
: > $out
export LD_LIBRARY_PATH=$L
for content in smooth noise dark; do
  for cfg in "100 60 17 10 3 3 1" "120 90 17 10 3 3 2" "400 133 17 10 3 3 1" "96 54 3 2 3 3 2" "64 40 3 1 3 3 2" "1600 240 17 10 3 3 1" "160 60 3 1 2 3 1" "240 64 4 3 3 4 2" "512 64 5 3 3 3 1"; do
    timeout 120 $K $cfg $content 1 0 generic >> $out 2>&1 || echo "   ^^^ rc=$?" >> $out
  done
done
timeout 300 $K 4000 2000 17 10 3 3 4 smooth 5 0 generic >> $out 2>&1
timeout 300 $K 4000 2000 17 10 3 3 4 noise 5 0 >> $out 2>&1
timeout 300 $K 4000 2000 17 10 3 3 4 smooth 5 8 >> $out 2>&1
timeout 300 $K 16384 4096 17 10 3 3 1 smooth 3 0 >> $out 2>&1
timeout 300 $K 1920 1080 3 1 3 3 8 smooth 5 0 generic >> $out 2>&1
grep -c " 0 bytes differ" $out
grep "differ" $out | grep -v " 0 bytes differ" | head
grep "frames=4 \|16384x4096\|1920x1080" $out | grep -v "^ " | cut -c1-200
