#!/bin/bash
TAG=r03b
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -8 > gpurun_out/${TAG}_tests.log
cat gpurun_out/${TAG}_tests.log
for v in g_over h_bg; do
for c in C2 C3; do
echo $v | tee -a gpurun_out/${TAG}_ab.log
FLUIDMARCH_LIB=$PWD/build_variants/$v/libfluidmarch.so FLUIDMARCH_AB=1 timeout 300 python tools/latency_probe.py $c 40 2>&1 | tail -1 | tee -a gpurun_out/${TAG}_ab.log
done
done
