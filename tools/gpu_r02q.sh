#!/bin/bash
TAG=r02q
mkdir -p gpurun_out
for occ in 0 1; do
  export FLUIDMARCH_OCC_SMEM=$occ
  echo "== OCC_SMEM=$occ" | tee -a gpurun_out/${TAG}_ab.log
  timeout 600 python tools/ab_probe.py C2 2>&1 | tee -a gpurun_out/${TAG}_ab.log
  timeout 600 python tools/ab_probe.py C3 2>&1 | tee -a gpurun_out/${TAG}_ab.log
done
