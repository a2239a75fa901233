#!/bin/bash
TAG=r03e
mkdir -p gpurun_out
for lanes in 7 8 9 5 7 8 9 5; do
timeout 600 python bench.py --lanes $lanes --no-tiles --no-aniso --no-cpu-baseline > gpurun_out/${TAG}_b.json 2> gpurun_out/${TAG}_bench.err
python - <<EOF2 | tee -a gpurun_out/${TAG}_lanes.log
import json
for l in open('gpurun_out/${TAG}_b.json'):
    if l.startswith('{'):
        d=json.loads(l); print('lanes $lanes', 'value ms', round(d['ms_per_step'],4), 'e2e ms', round(d['e2e']['ms_per_step'],4), 'present', round((d['e2e'].get('present') or {}).get('ms_per_step',0),4), 'refproto', round((d['e2e'].get('reference_protocol') or {}).get('ms_per_step',0),4),'ceiling', round(d['e2e']['host_ceiling']['ms_per_frame_pair'],4))
EOF2
done
