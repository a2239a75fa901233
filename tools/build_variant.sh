#!/bin/bash
# builds an A/B variant of libfluidmarch.so:  tools/build_variant.sh NAME "-DFM_UNROLL=2 -DFM_MINBLOCKS=4"
set -e
cd "$(dirname "$0")/.."
NAME=$1; EXTRA=$2
D=build_variants/$NAME; mkdir -p $D
for f in fm_context fm_grid fm_depth fm_march fm_query; do
  /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC $EXTRA -c bachelor-thesis_b200/csrc/$f.cu -o $D/$f.o &
done
wait
/usr/local/cuda/bin/nvcc -shared -o $D/libfluidmarch.so $D/*.o 2>/dev/null
rm -f $D/*.o
echo built $D/libfluidmarch.so
