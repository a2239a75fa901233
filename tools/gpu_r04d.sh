#!/bin/bash
TAG=r04d
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -4 | tee gpurun_out/${TAG}_tests.log
rm -rf build_variants/k_clear
for c in C2 C3; do timeout 600 python tools/ab_probe.py $c 2>&1 | cut -c1-200 | tee -a gpurun_out/${TAG}_ab.log; done
timeout 600 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none -k regex:k_key_count -c 3 --csv --log-file gpurun_out/${TAG}_key.csv python tools/prof_step.py C2 3 > /dev/null 2>&1
grep k_key_count gpurun_out/${TAG}_key.csv | tail -2 | cut -c1-50,200-400
timeout 600 python bench.py --steps 200 --no-tiles --no-aniso --no-cpu-baseline > gpurun_out/${TAG}_b.json 2> gpurun_out/${TAG}_bench.err
python - <<EOF2
import json
for l in open('gpurun_out/${TAG}_b.json'):
    if l.startswith('{'):
        d=json.loads(l); print('value ms', round(d['ms_per_step'],4), 'e2e ms', round(d['e2e']['ms_per_step'],4), 'lat', round(d['config'].get('latency_ms_per_frame'),4), 'grid', round(d['config']['stage_ms']['grid_ms'],4), d['parity']['pixels_differing'])
EOF2
