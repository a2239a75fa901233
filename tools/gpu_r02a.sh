#!/bin/bash
# round 2, first GPU call: parity (incl. C2 / C3 against oracle/_ref), A/B of the staged k_march_first, ncu capture
TAG=r02a
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -25 > gpurun_out/${TAG}_tests.log
cat gpurun_out/${TAG}_tests.log
timeout 600 python tools/ab_probe.py C2 > gpurun_out/${TAG}_ab_C2.log 2>&1
cat gpurun_out/${TAG}_ab_C2.log
timeout 300 python bench.py --steps 60 --warmup 12 --no-cpu-baseline > gpurun_out/${TAG}_bench_C2.json 2> gpurun_out/${TAG}_bench.err
cut -c1-300 gpurun_out/${TAG}_bench_C2.json; tail -3 gpurun_out/${TAG}_bench.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_march_first -s 2 -c 1 -o gpurun_out/${TAG}_march_first -f python tools/prof_step.py C2 3 > gpurun_out/${TAG}_ncu_march_first.log 2>&1
tail -2 gpurun_out/${TAG}_ncu_march_first.log
