#!/bin/bash
TAG=r02c
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -30 > gpurun_out/${TAG}_tests.log
cat gpurun_out/${TAG}_tests.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 600 python bench.py --steps 60 --warmup 12 > gpurun_out/${TAG}_bench_C2.json 2> gpurun_out/${TAG}_bench.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/r02c_bench_C2.json"))
print("ms/step", d["ms_per_step"], "latency", d["config"]["latency_ms_per_frame"], "e2e", d["e2e"]["ms_per_step"], "launches", d["gpu_launches"])
print(d["config"]["stage_ms"]); print({k: d[k] for k in ("parity_checked","pixels_differing")}, d["parity"]); print(d["cpu_baseline"])
r=d["roofline"]; print({k: r[k] for k in ("bound","achieved","peak","frac","traffic")})
PY
tail -5 gpurun_out/${TAG}_bench.err
timeout 300 python bench.py --impl reference --steps 5 --warmup 2 > gpurun_out/${TAG}_bench_ref.json 2>> gpurun_out/${TAG}_bench.err; cut -c1-600 gpurun_out/${TAG}_bench_ref.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches_C2.csv python tools/prof_step.py C2 3 > gpurun_out/${TAG}_prof.log 2>&1
tail -2 gpurun_out/${TAG}_prof.log
