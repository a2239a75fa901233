#!/bin/bash
# last captures of round 2: the three pre-pass kernels changed by r04p, the launch list, the bench line
TAG=r04r
mkdir -p gpurun_out
for k in depth_seed depth_bounds depth_splat; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_$k -s 2 -c 1 -o gpurun_out/${TAG}_$k -f python tools/prof_step.py C2 3 > gpurun_out/${TAG}_ncu_$k.log 2>&1
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches_C2.csv python tools/prof_step.py C2 3 > /dev/null 2>&1
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -3 | tee gpurun_out/${TAG}_tests.log
timeout 900 python bench.py > gpurun_out/${TAG}_bench_C2.json 2> gpurun_out/${TAG}_bench.err
tail -c 200 gpurun_out/${TAG}_bench.err
