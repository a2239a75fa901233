#!/bin/bash
TAG=r04m
mkdir -p gpurun_out
timeout 600 python tools/bg_check.py 2>&1 | tee gpurun_out/${TAG}_bgcheck.log
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -4 | tee gpurun_out/${TAG}_tests.log
for m in 0 1; do
FLUIDMARCH_BGFAST=$m timeout 300 python tools/prof_step.py C2 12 | cut -c1-230
FLUIDMARCH_BGFAST=$m timeout 300 python tools/prof_step.py C3 12 | cut -c1-230
done
timeout 600 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none -k regex:k_classify -c 3 --csv --log-file gpurun_out/${TAG}_cls.csv python tools/prof_step.py C2 3 > /dev/null 2>&1
grep "k_classify" gpurun_out/${TAG}_cls.csv | tail -2 | cut -c200-420
timeout 600 python bench.py --steps 200 --no-tiles --no-aniso --no-cpu-baseline > gpurun_out/${TAG}_b.json 2> gpurun_out/${TAG}_bench.err
python - <<EOF2
import json
for l in open('gpurun_out/${TAG}_b.json'):
    if l.startswith('{'):
        d=json.loads(l); print('value ms', round(d['ms_per_step'],4), 'e2e ms', round(d['e2e']['ms_per_step'],4), 'lat', round(d['config'].get('latency_ms_per_frame'),4))
EOF2
