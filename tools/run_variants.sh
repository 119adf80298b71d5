#!/bin/bash
# On the GPU box: time the decode of a few clips with every variant library, alternating (boxes differ by +-5 %, so only
# numbers of one run compare).  Usage: tools/run_variants.sh name1 name2 ...   (CLIPS="cfg frames;cfg frames" to override)
CLIPS=${CLIPS:-"cfg2_1080p_rgb32 200;cfg3_2160p_rgb32 30;cfg5_5120x1440 40;cfg4_1440p_intra 4"}
SCPR_LIB=variants/lib_$1.so timeout 300 python tools/stage_times.py cfg2_1080p_rgb32 60 > /dev/null 2>&1  # warm the box up
IFS=';' read -ra CL <<< "$CLIPS"
for rep in 1 2; do
  for c in "${CL[@]}"; do
    for v in "$@"; do
      echo -n "$v | $c | "; SCPR_LIB=variants/lib_$v.so timeout 300 python tools/stage_times.py $c 2>&1 | grep "rep 1" | sed 's/.*decode/decode/'
    done
  done
done
