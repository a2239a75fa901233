#!/bin/bash
TAG=r02n
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -5 > gpurun_out/${TAG}_tests.log
cat gpurun_out/${TAG}_tests.log
for pdl in 0 1; do
  export FLUIDMARCH_PDL=$pdl
  echo "== PDL=$pdl" | tee -a gpurun_out/${TAG}_ab.log
  timeout 600 python tools/ab_probe.py C2 2>&1 | tee -a gpurun_out/${TAG}_ab.log
  timeout 600 python tools/ab_probe.py C3 2>&1 | tee -a gpurun_out/${TAG}_ab.log
  for v in b_lean c_lean_mb2; do
    FLUIDMARCH_LIB=$PWD/build_variants/$v/libfluidmarch.so FLUIDMARCH_AB=1 timeout 300 python tools/latency_probe.py C2 40 2>&1 | tail -2 | tee -a gpurun_out/${TAG}_ab.log
  done
done
unset FLUIDMARCH_PDL
timeout 600 python bench.py --steps 20 --warmup 5 --no-tiles --no-aniso > gpurun_out/${TAG}_bench_C2.json 2> gpurun_out/${TAG}_bench.err
tail -c 1500 gpurun_out/${TAG}_bench_C2.json
