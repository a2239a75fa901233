#!/bin/bash
# 8-GPU box: frame-parallel + C3 region / tile-parallel records at N = 2, 4, 8
TAG=r04x
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8; nproc; free -g | head -2
timeout 600 python -m pytest tests/test_gpu_multi.py -q 2>&1 | tail -3 | tee gpurun_out/${TAG}_tests_multi.log
for N in 8 4 2; do
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500+N)) bench.py --gpus $N --steps 40 --warmup 8 > gpurun_out/${TAG}_bench_n$N.json 2> gpurun_out/${TAG}_bench_n$N.err
  tail -2 gpurun_out/${TAG}_bench_n$N.err | cut -c1-300
  python - <<PY
import json
try:
    d=json.loads([l for l in open("gpurun_out/${TAG}_bench_n$N.json") if l.startswith("{")][-1])
    print("N=$N ms/step", round(d["ms_per_step"],4), "value G", round(d["value"]/1e9,2), "e2e ms", round(d["e2e"]["ms_per_step"],4), "ceiling", {k: round(v,3) for k,v in d["e2e"]["host_ceiling"].items() if k!="note"}, "frac", round(d["e2e"]["fraction_of_host_ceiling"],3))
    t=d["tiles"]; print("  tiles:", {k: t[k] for k in ("ms_per_frame_one_gpu","ms_per_frame","ms_per_frame_wall_with_barrier","efficiency_vs_one_gpu","bit_identical","ms_per_rank","strips_px")}); print("  interleaved:", {k: t["interleaved_tiles"][k] for k in ("ms_per_frame","efficiency_vs_one_gpu","bit_identical")})
except Exception as e: print("N=$N failed", e)
PY
done
