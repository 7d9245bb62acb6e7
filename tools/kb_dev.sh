#!/bin/bash
# quick A/B of development variants (headline instance only): tools/kb_dev.sh VARIANT...   ("-" = product library)
K=tools/bin/kbench
for v in "$@"; do
  if [ "$v" = "-" ]; then L=lanczos_hls_b200; else L=variants/$v; fi
  echo "== $v"
  for c in smooth noise dark; do LD_LIBRARY_PATH=$L timeout 300 $K 1920 1080 2 1 3 3 64 $c 10 0 generic 2>&1 | grep -v "back to back\|streams" | cut -c60-330; done
  LD_LIBRARY_PATH=$L timeout 300 $K 1920 1080 2 1 3 3 64 smooth 10 8 2>&1 | grep -v "back to back\|2 stream\|4 stream" | cut -c1-330
  for cfg in "96 54 2 1 3 3 1" "1000 300 2 1 3 3 1" "960 540 2 1 3 3 3"; do
    for c in noise dark; do LD_LIBRARY_PATH=$L timeout 120 $K $cfg $c 1 0 generic 2>&1 | grep "vs generic" | cut -c1-200; done
  done
done
