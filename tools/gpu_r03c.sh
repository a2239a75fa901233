#!/bin/bash
TAG=r03c
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -4 > gpurun_out/${TAG}_tests.log
cat gpurun_out/${TAG}_tests.log
for lanes in 4 6 8 12; do
timeout 600 python bench.py --steps 200 --warmup 10 --lanes $lanes --no-tiles --no-aniso --no-cpu-baseline > gpurun_out/${TAG}_bench_lanes$lanes.json 2> gpurun_out/${TAG}_bench.err
python - <<EOF2
import json
for l in open('gpurun_out/${TAG}_bench_lanes$lanes.json'):
    if l.startswith('{'):
        d=json.loads(l); print('lanes $lanes', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'], 'present', (d['e2e'].get('present') or {}).get('ms_per_step'), 'lat', d['config'].get('latency_ms_per_frame'), d.get('parity',{}).get('pixels_differing'))
EOF2
done
