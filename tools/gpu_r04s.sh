#!/bin/bash
TAG=r04s
mkdir -p gpurun_out
for c in C1 C2 C3; do timeout 600 python tools/ab_probe.py $c 2>&1 | cut -c1-200 | tee -a gpurun_out/${TAG}_ab.log; done
