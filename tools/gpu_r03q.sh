#!/bin/bash
TAG=r03q
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -4 | tee gpurun_out/${TAG}_tests.log
CS=/usr/local/cuda/bin/compute-sanitizer
timeout 600 $CS --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_build.py -x -q -k "prepass_beside" 2>&1 | grep -E "passed|failed|SUMMARY" | tee -a gpurun_out/${TAG}_tests.log
timeout 600 $CS --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_build.py tests/test_cabi.py -m gpu -x -q 2>&1 | grep -E "passed|failed|SUMMARY" | tee -a gpurun_out/${TAG}_tests.log
timeout 600 python bench.py --steps 100 --no-tiles --no-aniso --no-cpu-baseline > gpurun_out/${TAG}_b.json 2> gpurun_out/${TAG}_bench.err
python - <<EOF2
import json
for l in open('gpurun_out/${TAG}_b.json'):
    if l.startswith('{'):
        d=json.loads(l); print('value ms', round(d['ms_per_step'],4), 'e2e ms', round(d['e2e']['ms_per_step'],4), 'lat', d['config'].get('latency_ms_per_frame'), d['config'].get('latency_with_stage_events_ms'), d.get('parity',{}).get('pixels_differing'))
EOF2
