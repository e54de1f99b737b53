#!/bin/bash
# Builds gala_b200/libgala_b200_<name>.so: the default objects with the listed kernel parts recompiled (fast build)
# with extra nvcc flags -- A/B experiments on the GPU box select it with GALA_B200_LIB=...
# usage: tools/build_variant.sh <name> "<extra nvcc flags>" [parts, default "2"]
set -e
NAME=$1; EXTRA=$2; PARTS=${3:-2}
cd "$(dirname "$0")/../gala_b200/csrc"
mkdir -p build_$NAME
OBJS=""
for f in build/*.o; do
  b=$(basename $f); skip=0
  for p in $PARTS; do [ "$b" = "kernels_fast_$p.o" ] && skip=1; done
  [ $skip = 0 ] && OBJS="$OBJS $f"
done
for p in $PARTS; do
  /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -Xptxas -v \
    $EXTRA -DGB_STRICT=0 -DGB_NS=gbk_fast -DGB_PART=$p -c kernels.cu -o build_$NAME/kernels_fast_$p.o 2> build_$NAME/ptxas_fast_$p.log &
done
wait
for p in $PARTS; do OBJS="$OBJS build_$NAME/kernels_fast_$p.o"; done
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../libgala_b200_$NAME.so $OBJS
grep -A2 "k_dop853_dynINS_9CompositeILi3EEELb0ELb1" build_$NAME/ptxas_fast_2.log 2>/dev/null | grep -E "registers|spill" | head -3
echo "built gala_b200/libgala_b200_$NAME.so"
