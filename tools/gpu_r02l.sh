#!/bin/bash
TAG=r02m
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -6 > gpurun_out/${TAG}_tests.log
cat gpurun_out/${TAG}_tests.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/${TAG}_bench_C2.json 2> gpurun_out/${TAG}_bench.err
tail -3 gpurun_out/${TAG}_bench.err
python - <<'PY'
import json
d=json.loads([l for l in open("gpurun_out/r02m_bench_C2.json") if l.startswith("{")][-1])
print("ms/step", d["ms_per_step"], "value G", d["value"]/1e9, "latency", d["config"]["latency_ms_per_frame"], "e2e", d["e2e"]["ms_per_step"], "launches", d["gpu_launches"])
print(d["config"]["stage_ms"]); print({k: d[k] for k in ("parity_checked","pixels_differing")}); print(d["cpu_baseline"]["value"], d["cpu_baseline"]["sample"][:120])
r=d["roofline"]; print({k: r[k] for k in ("bound","achieved","peak","frac","traffic")}, r["hbm_algorithmic"]["frac"])
print("e2e", {k: d["e2e"][k] for k in d["e2e"] if k not in ("host_ceiling","present","reference_protocol","path")}); print("present", d["e2e"].get("present"))
print("tiles", d["tiles"]); a=d["aniso"]; print("aniso", {k: a[k] for k in a if k!="workload"})
PY
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 | cut -c1-400
