#!/bin/bash
TAG=r03o
mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/${TAG}_bench_C2.json 2> gpurun_out/${TAG}_bench.err
tail -c 300 gpurun_out/${TAG}_bench.err
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | cut -c1-200
