#!/bin/bash
TAG=r04c
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none -c 200 --csv --log-file gpurun_out/${TAG}_launches_C3.csv python tools/prof_step.py C3 3 > gpurun_out/${TAG}_prof.log 2>&1
tail -1 gpurun_out/${TAG}_prof.log | cut -c1-200
