#!/bin/bash
TAG=r03p
mkdir -p gpurun_out
for bg in 0 1; do for c in C1 C2 C3; do
echo "BG=$bg" | tee -a gpurun_out/${TAG}_ab.log
FLUIDMARCH_BG=$bg timeout 300 python tools/latency_probe.py $c 40 2>&1 | tail -1 | tee -a gpurun_out/${TAG}_ab.log
done; done
FLUIDMARCH_BG=1 timeout 600 python -m pytest tests/test_gpu_build.py tests/test_cabi.py -m gpu -q -x 2>&1 | tail -3
