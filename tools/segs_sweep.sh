#!/bin/bash
for s in 0 3 5 6 9 13 18 27; do
  echo "== LZB_SEGS=$s"
  LZB_SEGS=$s LD_LIBRARY_PATH=lanczos_hls_b200 tools/bin/kbench 1920 1080 2 1 3 3 64 smooth 5 16 2>&1 | grep "single-frame\|frames=64" | cut -c1-140
done
