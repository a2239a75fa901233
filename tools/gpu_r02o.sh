#!/bin/bash
TAG=r02o
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_aniso.py -m gpu -q -x 2>&1 | tail -15 > gpurun_out/${TAG}_tests_aniso.log
cat gpurun_out/${TAG}_tests_aniso.log
bash tools/ab_aniso.sh 2>&1 | tee gpurun_out/${TAG}_ab_aniso.log
FLUIDMARCH_AB=1 FLUIDMARCH_LIB=$PWD/build_variants/zz_prof/libfluidmarch.so python tools/prof_step.py C2 6 2>&1 | tee gpurun_out/${TAG}_longprof.log
FLUIDMARCH_AB=1 FLUIDMARCH_LIB=$PWD/build_variants/zz_prof/libfluidmarch.so python tools/prof_step.py C3 6 2>&1 | tee -a gpurun_out/${TAG}_longprof.log
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -5 > gpurun_out/${TAG}_tests.log
cat gpurun_out/${TAG}_tests.log
