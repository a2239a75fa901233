#!/bin/bash
TAG=r05b
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_aniso.py -q -x 2>&1 | tail -3 | tee gpurun_out/${TAG}_tests.log
for m in 0 1 0 1; do
echo "ANISO_EMPTY=$m" | tee -a gpurun_out/${TAG}_ab.log
FLUIDMARCH_ANISO_EMPTY=$m timeout 300 python tools/prof_step_aniso.py C2 4 | cut -c1-330 | tee -a gpurun_out/${TAG}_ab.log
done
