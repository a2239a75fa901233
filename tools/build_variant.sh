#!/bin/bash
# builds an A/B variant of libfluidmarch.so:  tools/build_variant.sh NAME "-DFM_STAGE_CAP=640" [SRCDIR]
# SRCDIR (default: this tree's csrc) lets an older commit be built for comparison:
#   git worktree add /tmp/r01 <commit>; tools/build_variant.sh r01 "" /tmp/r01/bachelor-thesis_b200/csrc
set -e
cd "$(dirname "$0")/.."
NAME=$1; EXTRA=$2; SRC=${3:-bachelor-thesis_b200/csrc}
D=build_variants/$NAME; mkdir -p $D
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC $EXTRA"
for f in fm_context fm_grid fm_depth fm_march fm_query fm_sequence fm_bgeo fm_record fm_smooth; do
  /usr/local/cuda/bin/nvcc $FLAGS -c $SRC/$f.cu -o $D/$f.o &
done
/usr/local/cuda/bin/nvcc $FLAGS -fmad=false -DFM_NO_FMAD -c $SRC/fm_aniso.cu -o $D/fm_aniso.o &
wait
/usr/local/cuda/bin/nvcc -shared -o $D/libfluidmarch.so $D/*.o -lpthread -lz
rm -f $D/*.o
echo built $D/libfluidmarch.so
