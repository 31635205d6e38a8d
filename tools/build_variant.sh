#!/bin/bash
# Build an A/B variant of libb200gs.so (dev tool): the translation units named in FILES are recompiled with extra -D
# flags, everything else is taken from build/obj.   usage: tools/build_variant.sh TAG "FILES" "-DX=1 -DY=2"
# e.g. tools/build_variant.sh nostage "project" "-DPROJECT_BWD_STAGE_SH=0"  ->  build/ab/nostage.so
set -e
TAG=$1; FILES=$2; DEFS=$3
ROOT=$(cd "$(dirname "$0")/.." && pwd)
mkdir -p $ROOT/build/ab/obj_$TAG
OBJS=""
for o in $ROOT/build/obj/*.o; do
  b=$(basename $o .o)
  if [[ " $FILES " == *" $b "* ]]; then
    nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -diag-suppress 177 $DEFS \
      -c -o $ROOT/build/ab/obj_$TAG/$b.o $ROOT/robosimgs_b200/csrc/$b.cu &
    OBJS="$OBJS $ROOT/build/ab/obj_$TAG/$b.o"
  else
    OBJS="$OBJS $o"
  fi
done
wait
nvcc -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC -shared -o $ROOT/build/ab/$TAG.so $OBJS -lcudart
echo built $ROOT/build/ab/$TAG.so
