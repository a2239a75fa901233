#!/bin/bash
# pipelined bench (value / e2e) for every libfluidmarch.so under build_variants/
cd "$(dirname "$0")/.."
for d in build_variants/*/; do
  n=$(basename $d)
  FLUIDMARCH_AB=1 FLUIDMARCH_LIB=$PWD/$d/libfluidmarch.so timeout 300 python bench.py --steps ${STEPS:-60} --warmup 12 --no-cpu-baseline ${BENCH_ARGS} > gpurun_out/ab_$n.json 2> gpurun_out/ab_$n.err || tail -3 gpurun_out/ab_$n.err
  python - <<PY
import json
d=json.load(open("gpurun_out/ab_$n.json"))
print("$n", "ms/step", round(d["ms_per_step"],4), "latency", round(d["config"]["latency_ms_per_frame"],4), "e2e", round(d["e2e"]["ms_per_step"],4), {k: round(v,4) for k,v in d["config"]["stage_ms"].items() if v})
PY
done
