#!/bin/bash
TAG=r04n
mkdir -p gpurun_out
for k in classify march_first march_long key_count; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_$k -s 2 -c 1 -o gpurun_out/${TAG}_$k -f python tools/prof_step.py C2 3 > gpurun_out/${TAG}_ncu_$k.log 2>&1
done
ls gpurun_out/${TAG}_*.ncu-rep | wc -l
