#!/bin/bash
# one ncu --set full capture of the headline kernel per library variant (image-like content unless CONTENT is set):
#   tools/ncu_ab.sh VARIANT...   ("-" = product library) -> gpurun_out/prof_ab_<variant>.ncu-rep
K=tools/bin/kbench
for v in "$@"; do
  if [ "$v" = "-" ]; then L=lanczos_hls_b200; n=product; else L=variants/$v; n=$v; fi
  LD_LIBRARY_PATH=$L $K 1920 1080 2 1 3 3 64 ${CONTENT:-smooth} 1 0 > /dev/null 2>&1 && \
  LD_LIBRARY_PATH=$L ncu --set full --clock-control none --import-source on -k regex:lanczos_v6 -s 3 -c 1 -f -o gpurun_out/prof_ab_$n \
      $K 1920 1080 2 1 3 3 64 ${CONTENT:-smooth} 1 0 > gpurun_out/ncu_ab_$n.log 2>&1
done
ls -la gpurun_out/prof_ab_*
