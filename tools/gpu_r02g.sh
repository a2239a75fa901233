#!/bin/bash
TAG=r02g
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -15 > gpurun_out/${TAG}_tests.log
cat gpurun_out/${TAG}_tests.log
timeout 300 python tools/prof_step.py C2 12 | cut -c1-330
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 40 --warmup 8 > gpurun_out/${TAG}_bench_n2.json 2> gpurun_out/${TAG}_bench_n2.err
tail -3 gpurun_out/${TAG}_bench_n2.err
python - <<'PY'
import json
d=json.loads([l for l in open("gpurun_out/r02g_bench_n2.json") if l.startswith("{")][-1])
print("N=2 ms/step", d["ms_per_step"], "value", d["value"], "e2e", d["e2e"]["ms_per_step"], d["e2e"].get("host_ceiling"), d["e2e"].get("fraction_of_host_ceiling"))
print("tiles", d["tiles"])
PY
