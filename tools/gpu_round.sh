#!/bin/bash
# one GPU round: parity tests, bench line, ncu launch list, full captures of the top kernels.  usage: tools/gpu_round.sh r01s
TAG=${1:-rXX}
set -x
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/${TAG}_tests.log
cat gpurun_out/${TAG}_tests.log
timeout 600 python bench.py > gpurun_out/${TAG}_bench_C2.json 2> gpurun_out/${TAG}_bench.err
cut -c1-400 gpurun_out/${TAG}_bench_C2.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches_C2.csv python tools/prof_step.py C2 3 > gpurun_out/${TAG}_prof.log 2>&1
tail -2 gpurun_out/${TAG}_prof.log
for k in march_first depth_splat march_long; do
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_$k -s 2 -c 1 -o gpurun_out/${TAG}_$k -f python tools/prof_step.py C2 3 > gpurun_out/${TAG}_ncu_$k.log 2>&1
done
ls -la gpurun_out | tail -12
