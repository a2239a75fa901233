#!/bin/bash
TAG=r02e
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -30 > gpurun_out/${TAG}_tests.log
cat gpurun_out/${TAG}_tests.log
echo "--- async (view from device memory)"; timeout 300 python tools/prof_step.py C2 12 | cut -c1-330
echo "--- sync (view by value)"; FR_ASYNC=0 timeout 300 python tools/prof_step.py C2 12 | cut -c1-330
echo "--- no depth refine"; FR_DEPTH_REFINE=0 timeout 300 python tools/prof_step.py C2 12 | cut -c1-330
timeout 600 python bench.py --steps 60 --warmup 12 --no-cpu-baseline > gpurun_out/${TAG}_bench_C2.json 2> gpurun_out/${TAG}_bench.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/r02e_bench_C2.json"))
print("ms/step", d["ms_per_step"], "latency", d["config"]["latency_ms_per_frame"], "e2e", d["e2e"]["ms_per_step"], "launches", d["gpu_launches"])
print(d["config"]["stage_ms"])
PY
tail -5 gpurun_out/${TAG}_bench.err
timeout 300 python tools/prof_step.py C3 8 | cut -c1-330
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches_C2.csv python tools/prof_step.py C2 3 > gpurun_out/${TAG}_prof.log 2>&1
