#!/bin/bash
# Build a tuning variant of libttlda.so for same-box A/B runs:  tools/build_variant.sh NAME "-DFOO=1 -DBAR=2" [file.cu ...]
# Only the listed sources (default: estep.cu) are recompiled with the extra flags; the other objects come from
# tuotuo_b200/_obj (run `python -m tuotuo_b200.build` first).  Output: tools/lib/libttlda_NAME.so (git-ignored, shipped by
# gpurun); select it with TTLDA_LIB=tools/lib/libttlda_NAME.so.
set -e
NAME=$1; FLAGS=$2; shift 2 || true
SRCS=${@:-estep.cu}
ROOT=$(cd "$(dirname "$0")/.." && pwd)
mkdir -p $ROOT/tools/lib/obj_$NAME
OBJS=""
for o in $ROOT/tuotuo_b200/_obj/*.o; do
  b=$(basename $o .o)
  if echo " $SRCS " | grep -q " $b.cu "; then
    nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -I $ROOT/include -I $ROOT/tuotuo_b200/csrc \
      $FLAGS -c $ROOT/tuotuo_b200/csrc/$b.cu -o $ROOT/tools/lib/obj_$NAME/$b.o
    OBJS="$OBJS $ROOT/tools/lib/obj_$NAME/$b.o"
  else
    OBJS="$OBJS $o"
  fi
done
nvcc -shared -gencode arch=compute_100a,code=sm_100a -o $ROOT/tools/lib/libttlda_$NAME.so $OBJS
rm -rf $ROOT/tools/lib/obj_$NAME
ls -la $ROOT/tools/lib/libttlda_$NAME.so
