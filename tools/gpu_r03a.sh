#!/bin/bash
TAG=r03a
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -8 > gpurun_out/${TAG}_tests.log
cat gpurun_out/${TAG}_tests.log
for ov in 1 0; do
export FLUIDMARCH_OVERLAP=$ov
echo "== OVERLAP=$ov" | tee -a gpurun_out/${TAG}_ab.log
for c in C1 C2 C3; do
FLUIDMARCH_LIB=$PWD/build_variants/g_over/libfluidmarch.so FLUIDMARCH_AB=1 timeout 300 python tools/latency_probe.py $c 40 2>&1 | tail -1 | tee -a gpurun_out/${TAG}_ab.log
done
done
unset FLUIDMARCH_OVERLAP
FLUIDMARCH_LIB=$PWD/build_variants/a_head/libfluidmarch.so FLUIDMARCH_AB=1 timeout 300 python tools/latency_probe.py C2 40 2>&1 | tail -1 | tee -a gpurun_out/${TAG}_ab.log
for ovl in 0 1; do
FLUIDMARCH_OVERLAP_LANES=$ovl timeout 600 python bench.py --steps 100 --warmup 10 --no-tiles --no-aniso --no-cpu-baseline > gpurun_out/${TAG}_bench_lanes$ovl.json 2> gpurun_out/${TAG}_bench.err
python - <<EOF2
import json
for l in open('gpurun_out/${TAG}_bench_lanes$ovl.json'):
    if l.startswith('{'):
        d=json.loads(l); print('lanes_overlap$ovl', d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'], d['config'].get('latency_ms_per_frame'), d['config'].get('latency_with_stage_events_ms'), d.get('parity',{}).get('pixels_differing'))
EOF2
done
