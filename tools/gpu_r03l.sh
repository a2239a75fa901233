#!/bin/bash
TAG=r03l
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -6 > gpurun_out/${TAG}_tests.log
cat gpurun_out/${TAG}_tests.log
for c in C1 C2 C3; do timeout 600 python tools/ab_probe.py $c 2>&1 | tee -a gpurun_out/${TAG}_ab.log; done
bash tools/ab_aniso.sh 2>&1 | tee -a gpurun_out/${TAG}_ab.log
for v in a_head i_run; do for c in C2 C3; do
FLUIDMARCH_LIB=$PWD/build_variants/$v/libfluidmarch.so FLUIDMARCH_AB=1 timeout 300 python tools/latency_probe.py $c 40 2>&1 | tail -1 | tee -a gpurun_out/${TAG}_ab.log
done; done
