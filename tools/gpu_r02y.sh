#!/bin/bash
TAG=r02y
mkdir -p gpurun_out
for fuse in 1 0; do
export FLUIDMARCH_FUSE=$fuse
echo "== FUSE=$fuse" | tee -a gpurun_out/${TAG}_ab.log
timeout 600 python tools/ab_probe.py C2 2>&1 | tee -a gpurun_out/${TAG}_ab.log
timeout 600 python tools/ab_probe.py C3 2>&1 | tee -a gpurun_out/${TAG}_ab.log
done
unset FLUIDMARCH_FUSE
for c in C2 C3; do
FLUIDMARCH_LIB=$PWD/build_variants/zz_fprof/libfluidmarch.so FLUIDMARCH_AB=1 timeout 300 python tools/first_profile.py $c 2>&1 | tee -a gpurun_out/${TAG}_fprof.log
done
for v in a_head f_fuse; do
FLUIDMARCH_LIB=$PWD/build_variants/$v/libfluidmarch.so FLUIDMARCH_AB=1 timeout 300 python tools/latency_probe.py C2 40 2>&1 | tail -2 | tee -a gpurun_out/${TAG}_ab.log
done
