#!/bin/bash
TAG=r02i
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -15 > gpurun_out/${TAG}_tests.log
cat gpurun_out/${TAG}_tests.log
timeout 300 python tools/region_probe.py 8 2>&1 | cut -c1-300
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 40 --warmup 8 > gpurun_out/${TAG}_bench_n2.json 2> gpurun_out/${TAG}_bench_n2.err
tail -3 gpurun_out/${TAG}_bench_n2.err | cut -c1-300
python - <<'PY'
import json
d=json.loads([l for l in open("gpurun_out/r02i_bench_n2.json") if l.startswith("{")][-1])
print("N=2 ms/step", d["ms_per_step"], "e2e", d["e2e"]["ms_per_step"], "present", d["e2e"].get("present"))
t=d["tiles"]; print("tiles", {k: t[k] for k in ("ms_per_frame_one_gpu","ms_per_frame","efficiency_vs_one_gpu","bit_identical","ms_per_rank","strips_px")})
PY
