#!/bin/bash
# Timing ablations of the headline kernel (development): builds variants with parts of the kernel compiled out
# (results are WRONG, only the time is of interest) and prints the time of config 2, 64 frames, smooth content.
#   tools/ablate.sh build            (here, no GPU)      tools/ablate.sh run   (on the GPU box)
V="base: novfilter:-DLZB_ABL_NOVFILTER novguard:-DLZB_ABL_NOVGUARD nohfilter:-DLZB_ABL_NOHFILTER nov:-DLZB_ABL_NOV noh:-DLZB_ABL_NOH nostore:-DLZB_ABL_NOSTORE novf_novg:-DLZB_ABL_NOVFILTER,-DLZB_ABL_NOVGUARD"
if [ "$1" = "build" ]; then
  for v in $V; do n=${v%%:*}; f=$(echo ${v#*:} | tr ',' ' '); tools/build_variant.sh abl_$n -DLZB_V6_DEV $f > /dev/null || echo "build $n failed"; done
else
  for v in $V; do n=${v%%:*}
    for fl in 0 8; do
      LD_LIBRARY_PATH=variants/abl_$n tools/bin/kbench 1920 1080 2 1 3 3 64 smooth 10 $fl 2>&1 | grep "frames=64" | sed "s/^/$n /" | cut -c1-40,100-200
    done
  done
fi
