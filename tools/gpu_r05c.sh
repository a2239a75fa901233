#!/bin/bash
TAG=r05c
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -3 | tee gpurun_out/${TAG}_tests.log
timeout 900 python bench.py --no-tiles > gpurun_out/${TAG}_bench_C2.json 2> gpurun_out/${TAG}_bench.err
tail -c 200 gpurun_out/${TAG}_bench.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_march_long -s 1 -c 1 -o gpurun_out/${TAG}_aniso_march_long -f python tools/prof_step_aniso.py C2 2 > gpurun_out/${TAG}_ncu_aniso_march_long.log 2>&1
