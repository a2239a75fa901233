#!/bin/bash
TAG=r02b
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x 2>&1 | tail -5 > gpurun_out/${TAG}_tests.log
cat gpurun_out/${TAG}_tests.log
timeout 600 python tools/ab_probe.py C2 > gpurun_out/${TAG}_ab_C2.log 2>&1
cat gpurun_out/${TAG}_ab_C2.log
timeout 600 python tools/ab_probe.py C3 > gpurun_out/${TAG}_ab_C3.log 2>&1
cat gpurun_out/${TAG}_ab_C3.log
STEPS=60 BENCH_ARGS="" bash tools/ab_bench.sh 2>&1 | tail -8
