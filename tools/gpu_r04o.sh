#!/bin/bash
# end of round 2: tests, smoke, the default bench line and the reference arm on the committed code
TAG=r04o
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -3 | tee gpurun_out/${TAG}_tests.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | cut -c1-160
timeout 900 python bench.py > gpurun_out/${TAG}_bench_C2.json 2> gpurun_out/${TAG}_bench.err
tail -c 300 gpurun_out/${TAG}_bench.err
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${TAG}_bench_reference.json 2>> gpurun_out/${TAG}_bench.err
cut -c1-200 gpurun_out/${TAG}_bench_reference.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches_C2.csv python tools/prof_step.py C2 3 > /dev/null 2>&1
