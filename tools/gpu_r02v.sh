#!/bin/bash
TAG=r02v
mkdir -p gpurun_out
for k in march_first march_long; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_$k -s 2 -c 1 -o gpurun_out/${TAG}_$k -f python tools/prof_step.py C2 3 > gpurun_out/${TAG}_ncu_$k.log 2>&1
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_march_long -s 2 -c 1 -o gpurun_out/${TAG}_C3_march_long -f python tools/prof_step.py C3 3 > gpurun_out/${TAG}_ncu_C3_march_long.log 2>&1
ls -la gpurun_out/${TAG}_*
