#!/bin/bash
# A/B: tools/kb_ab.sh <variant dir|-> ... ; runs config 2 smooth exact + tolerance for each library variant
K=tools/bin/kbench
for v in "$@"; do
  if [ "$v" = "-" ]; then L=lanczos_hls_b200; else L=variants/$v; fi
  echo "== $v"
  LD_LIBRARY_PATH=$L timeout 300 $K 1920 1080 2 1 3 3 32 smooth 20 0 | cut -c60-200
  LD_LIBRARY_PATH=$L timeout 300 $K 1920 1080 2 1 3 3 32 smooth 20 8 | cut -c60-200
  LD_LIBRARY_PATH=$L timeout 300 $K 2560 1440 3 2 3 4 16 smooth 20 0 | cut -c60-200
done
