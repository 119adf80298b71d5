#!/bin/bash
# A/B two builds of the library on the SAME box (boxes differ by several percent): tools/ab.sh <variant.patch> [clip frames]...
# Builds the tree as it is (A), applies the patch and builds again (B), then times both alternately.
patch=$1; shift
make -s -C screenpressor_b200/csrc 2>&1 | grep -v deprecated; cp screenpressor_b200/libscpr_b200.so /tmp/libA.so
patch -p1 -s < $patch && make -s -C screenpressor_b200/csrc 2>&1 | grep -v deprecated; cp screenpressor_b200/libscpr_b200.so /tmp/libB.so
SCPR_LIB=/tmp/libA.so timeout 300 python tools/stage_times.py ${1:-cfg2_1080p_rgb32} ${2:-600} > /dev/null 2>&1  # warm the box up
for rep in 1 2 3; do
  for v in A B; do
    echo -n "$v: "; SCPR_LIB=/tmp/lib$v.so timeout 300 python tools/stage_times.py ${1:-cfg2_1080p_rgb32} ${2:-600} 2>&1 | grep "rep 1" | sed 's/.*decode/decode/'
  done
done
if [ -n "$3" ]; then for v in A B; do echo -n "$v $3: "; SCPR_LIB=/tmp/lib$v.so timeout 300 python tools/stage_times.py $3 $4 2>&1 | grep "rep 1" | sed 's/.*decode/decode/'; done; fi
SCPR_LIB=/tmp/libB.so timeout 600 python tools/dev_check.py fuzz cfg2 cfg1 2>&1 | tail -1
