#!/bin/bash
TAG=r03r
mkdir -p gpurun_out
: > gpurun_out/${TAG}_configs.jsonl
for c in c3 c4 c5 aniso; do
  timeout 600 python tools/bench_configs.py $c >> gpurun_out/${TAG}_configs.jsonl 2>> gpurun_out/${TAG}_configs.err
done
wc -l gpurun_out/${TAG}_configs.jsonl; tail -3 gpurun_out/${TAG}_configs.err
