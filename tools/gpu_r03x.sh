#!/bin/bash
TAG=r03x
mkdir -p gpurun_out
for cps in 3 2 3 2; do
FR_MARCH_CTAS_PER_SM=$cps timeout 600 python bench.py --steps 200 --no-tiles --no-aniso --no-cpu-baseline > gpurun_out/${TAG}_b.json 2> gpurun_out/${TAG}_bench.err
python - <<EOF2 | tee -a gpurun_out/${TAG}_ab.log
import json
for l in open('gpurun_out/${TAG}_b.json'):
    if l.startswith('{'):
        d=json.loads(l); print('CTAS_PER_SM=$cps value ms', round(d['ms_per_step'],4), 'e2e ms', round(d['e2e']['ms_per_step'],4), 'lat', round(d['config'].get('latency_ms_per_frame'),4), 'first', round(d['config']['stage_ms']['march_first_ms'],4))
EOF2
done
