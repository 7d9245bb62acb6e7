#!/bin/bash
K=tools/bin/kbench
for content in smooth noise dark; do
  for cfg in "100 60 17 10 3 3 1" "120 90 17 10 3 3 2" "400 133 17 10 3 3 1" "96 54 3 2 3 3 2" "64 40 3 1 3 3 2" "1600 240 17 10 3 3 1"; do
    timeout 120 $K $cfg $content 1 0 v5 2>&1 | cut -c1-330 || echo "   ^^^ rc=$?"
  done
done
timeout 300 $K 4000 2000 17 10 3 3 4 smooth 5 0 v5 2>&1 | cut -c1-330
timeout 300 $K 4000 2000 17 10 3 3 4 smooth 5 8 2>&1 | cut -c1-330
timeout 300 $K 16384 4096 17 10 3 3 1 smooth 3 0 2>&1 | cut -c1-330
