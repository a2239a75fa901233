#!/bin/bash
TAG=r03j
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_build.py -q -x 2>&1 | tail -25 > gpurun_out/${TAG}_tests_build.log
cat gpurun_out/${TAG}_tests_build.log
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -6 > gpurun_out/${TAG}_tests.log
cat gpurun_out/${TAG}_tests.log
timeout 300 python tools/prof_step.py C2 12 | cut -c1-260
timeout 300 python tools/prof_step.py C3 12 | cut -c1-260
