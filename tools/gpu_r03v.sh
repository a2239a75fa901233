#!/bin/bash
TAG=r03v
mkdir -p gpurun_out
for sh in 0 1; do for ov in 1 0; do for c in C2 C3; do
echo "SHUFFLE=$sh OVERLAP=$ov" | tee -a gpurun_out/${TAG}_ab.log
SHUFFLE=$sh FLUIDMARCH_OVERLAP=$ov timeout 300 python tools/latency_probe.py $c 30 2>&1 | tail -1 | tee -a gpurun_out/${TAG}_ab.log
done; done; done
