#!/bin/bash
# final 2-GPU check of the round: multi-GPU tests + the N = 2 bench line (frame-parallel, C3 regions / tiles)
TAG=r04w
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_multi.py -q 2>&1 | tail -3 | tee gpurun_out/${TAG}_tests_multi.log
N=2
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 60 --warmup 8 > gpurun_out/${TAG}_bench_n$N.json 2> gpurun_out/${TAG}_bench_n$N.err
tail -2 gpurun_out/${TAG}_bench_n$N.err | cut -c1-300
python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/${TAG}_bench_n$N.json") if l.startswith("{")][-1])
print("N=$N ms/step", round(d["ms_per_step"],4), "value G", round(d["value"]/1e9,2), "e2e ms", round(d["e2e"]["ms_per_step"],4), "frac", round(d["e2e"]["fraction_of_host_ceiling"],3), "present", d["e2e"]["present"]["ms_per_step"])
t=d["tiles"]; print("  tiles:", {k: t[k] for k in ("ms_per_frame_one_gpu","ms_per_frame","efficiency_vs_one_gpu","bit_identical","ms_per_rank")}); print("  interleaved:", {k: t["interleaved_tiles"][k] for k in ("ms_per_frame","efficiency_vs_one_gpu","bit_identical")})
PY
