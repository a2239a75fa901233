#!/bin/bash
# final profile round of the round-2 code (tag r04h): tests, launch list, ncu --set full of every kernel, bench line, reference arm
TAG=r04h
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -4 > gpurun_out/${TAG}_tests.log
cat gpurun_out/${TAG}_tests.log
bash tools/gpu_profile.sh ${TAG}
timeout 900 python bench.py > gpurun_out/${TAG}_bench_C2.json 2> gpurun_out/${TAG}_bench.err
tail -c 300 gpurun_out/${TAG}_bench.err
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${TAG}_bench_reference.json 2>> gpurun_out/${TAG}_bench.err
cut -c1-200 gpurun_out/${TAG}_bench_reference.json
