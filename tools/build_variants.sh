#!/bin/bash
# Build several variants of libscpr_b200.so (compile-time switches of csrc/decode.cu) into variants/ for a same-box A/B run:
#   tools/build_variants.sh name1 "-DSCPR_X=1 ..." name2 "..." ...      (run here; the .so files travel with gpurun)
#   "old:<git rev>" as the flags builds that revision's decode.cu instead.
# then on the GPU box: tools/run_variants.sh name1 name2 ...
set -e
cd "$(dirname "$0")/.."
CS=screenpressor_b200/csrc
NVF="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC"
make -s -C $CS 2>&1 | grep -v deprecated || true
mkdir -p variants
while [ $# -gt 1 ]; do
  name=$1; flags=$2; shift 2
  src=$CS/decode.cu
  if [[ "$flags" == old:* ]]; then git show "${flags#old:}:$CS/decode.cu" > $CS/decode_variant_tmp.cu; src=$CS/decode_variant_tmp.cu; flags=""; fi
  nvcc $NVF $flags -c $src -o /tmp/decode_$name.o
  objs=$(ls $CS/build/*.o | grep -v decode.o)
  nvcc -shared -o variants/lib_$name.so $objs /tmp/decode_$name.o -lcudart 2>&1 | grep -v deprecated || true
  rm -f $CS/decode_variant_tmp.cu
  echo "built variants/lib_$name.so ($flags)"
done
