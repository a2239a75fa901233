#!/bin/bash
TAG=r03m
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -6 > gpurun_out/${TAG}_tests.log
cat gpurun_out/${TAG}_tests.log
timeout 300 python tools/prof_step.py C2 12 | cut -c1-260
timeout 300 python tools/prof_step_aniso.py C2 4 | cut -c1-330
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_march_long -s 2 -c 1 -o gpurun_out/${TAG}_march_long -f python tools/prof_step.py C2 3 > gpurun_out/${TAG}_ncu_march_long.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches_C2.csv python tools/prof_step.py C2 3 > gpurun_out/${TAG}_prof.log 2>&1
