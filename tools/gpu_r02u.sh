#!/bin/bash
TAG=r02u
mkdir -p gpurun_out
for c in C2 C3; do
FLUIDMARCH_LIB=$PWD/build_variants/zz_fprof/libfluidmarch.so FLUIDMARCH_AB=1 timeout 300 python tools/first_profile.py $c 2>&1 | tee -a gpurun_out/${TAG}_fprof.log
done
