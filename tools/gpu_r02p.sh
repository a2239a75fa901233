#!/bin/bash
TAG=r02p
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -5 > gpurun_out/${TAG}_tests.log
cat gpurun_out/${TAG}_tests.log
for occ in 0 1; do
  export FLUIDMARCH_OCC_SMEM=$occ
  echo "== OCC_SMEM=$occ" | tee -a gpurun_out/${TAG}_ab.log
  timeout 600 python tools/ab_probe.py C2 2>&1 | tee -a gpurun_out/${TAG}_ab.log
  timeout 600 python tools/ab_probe.py C3 2>&1 | tee -a gpurun_out/${TAG}_ab.log
  bash tools/ab_aniso.sh 2>&1 | tee -a gpurun_out/${TAG}_ab.log
done
unset FLUIDMARCH_OCC_SMEM
FLUIDMARCH_LIB=$PWD/build_variants/b_new/libfluidmarch.so FLUIDMARCH_AB=1 timeout 300 python tools/latency_probe.py C2 40 2>&1 | tail -2 | tee -a gpurun_out/${TAG}_ab.log
FLUIDMARCH_LIB=$PWD/build_variants/b_new/libfluidmarch.so FLUIDMARCH_AB=1 timeout 300 python tools/latency_probe.py C3 20 2>&1 | tail -2 | tee -a gpurun_out/${TAG}_ab.log
for pdl in 0 1; do
FLUIDMARCH_PDL=$pdl timeout 600 python bench.py --steps 60 --warmup 10 --no-tiles --no-aniso > gpurun_out/${TAG}_bench_pdl$pdl.json 2> gpurun_out/${TAG}_bench.err
python - <<EOF
import json
for l in open('gpurun_out/${TAG}_bench_pdl$pdl.json'):
    if l.startswith('{'):
        d=json.loads(l); print('pdl$pdl', d['value'], d['ms_per_step'], d['e2e']['value'], d['config'].get('latency_ms_per_frame'), d['config'].get('latency_with_stage_events_ms'), d.get('parity',{}).get('pixels_differing'))
EOF
done
