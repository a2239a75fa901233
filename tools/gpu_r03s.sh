#!/bin/bash
TAG=r03s
mkdir -p gpurun_out
for bg in 0 1 2; do for c in C1 C2 C3; do
echo "BG=$bg" | tee -a gpurun_out/${TAG}_ab.log
FLUIDMARCH_BG=$bg timeout 300 python tools/latency_probe.py $c 40 2>&1 | tail -1 | tee -a gpurun_out/${TAG}_ab.log
done; done
for bg in 1 2 1 2; do
FLUIDMARCH_BG=$bg timeout 600 python bench.py --steps 200 --no-tiles --no-aniso --no-cpu-baseline > gpurun_out/${TAG}_b.json 2> gpurun_out/${TAG}_bench.err
python - <<EOF2 | tee -a gpurun_out/${TAG}_ab.log
import json
for l in open('gpurun_out/${TAG}_b.json'):
    if l.startswith('{'):
        d=json.loads(l); print('BG=$bg value ms', round(d['ms_per_step'],4), 'e2e ms', round(d['e2e']['ms_per_step'],4), 'lat', round(d['config'].get('latency_ms_per_frame'),4), round(d['config'].get('latency_with_stage_events_ms'),4), d['config']['stage_ms']['classify_ms'], d['config']['stage_ms']['march_first_ms'], d.get('parity',{}).get('pixels_differing'))
EOF2
done
FLUIDMARCH_BG=2 timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
