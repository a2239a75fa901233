#!/bin/bash
TAG=r02z
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -8 > gpurun_out/${TAG}_tests.log
cat gpurun_out/${TAG}_tests.log
for trim in 1 0; do
export FLUIDMARCH_TRIM=$trim
echo "== TRIM=$trim" | tee -a gpurun_out/${TAG}_ab.log
timeout 600 python tools/ab_probe.py C2 2>&1 | tee -a gpurun_out/${TAG}_ab.log
timeout 600 python tools/ab_probe.py C3 2>&1 | tee -a gpurun_out/${TAG}_ab.log
timeout 600 python tools/ab_probe.py C1 2>&1 | tee -a gpurun_out/${TAG}_ab.log
FLUIDMARCH_AB=1 FLUIDMARCH_LIB=$PWD/build_variants/f_trim/libfluidmarch.so python tools/prof_step_aniso.py C2 4 | cut -c1-330 | tee -a gpurun_out/${TAG}_ab.log
done
unset FLUIDMARCH_TRIM
for c in C2 C3; do
FLUIDMARCH_LIB=$PWD/build_variants/zz_fprof/libfluidmarch.so FLUIDMARCH_AB=1 timeout 300 python tools/first_profile.py $c 2>&1 | tee -a gpurun_out/${TAG}_fprof.log
done
for v in a_head f_trim; do
FLUIDMARCH_LIB=$PWD/build_variants/$v/libfluidmarch.so FLUIDMARCH_AB=1 timeout 300 python tools/latency_probe.py C2 40 2>&1 | tail -2 | tee -a gpurun_out/${TAG}_ab.log
done
