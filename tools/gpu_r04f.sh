#!/bin/bash
TAG=r04f
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -4 | tee gpurun_out/${TAG}_tests.log
for c in C1 C2 C3; do timeout 600 python tools/ab_probe.py $c 2>&1 | cut -c1-200 | tee -a gpurun_out/${TAG}_ab.log; done
for v in a_head n_seed; do for c in C2 C3; do
echo $v | tee -a gpurun_out/${TAG}_ab.log
FLUIDMARCH_LIB=$PWD/build_variants/$v/libfluidmarch.so FLUIDMARCH_AB=1 timeout 300 python tools/latency_probe.py $c 40 2>&1 | tail -1 | tee -a gpurun_out/${TAG}_ab.log
done; done
timeout 600 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none -k regex:k_depth_seed -c 6 --csv --log-file gpurun_out/${TAG}_cull.csv python tools/prof_step.py C2 3 > /dev/null 2>&1
grep "k_depth_seed" gpurun_out/${TAG}_cull.csv | tail -4 | cut -c60-110,200-400
