set -x
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r01j_tests.log
cat gpurun_out/r01j_tests.log
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/r01j_bench_C2.json 2> gpurun_out/r01j_bench.err
cat gpurun_out/r01j_bench_C2.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r01j_launches_C2.csv python tools/prof_step.py C2 3 > gpurun_out/r01j_prof.log 2>&1
tail -3 gpurun_out/r01j_prof.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_march_first -s 2 -c 1 -o gpurun_out/r01j_march_first -f python tools/prof_step.py C2 3 > gpurun_out/r01j_ncu1.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_depth_splat -s 2 -c 1 -o gpurun_out/r01j_depth_splat -f python tools/prof_step.py C2 3 > gpurun_out/r01j_ncu2.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_march_long -s 2 -c 1 -o gpurun_out/r01j_march_long -f python tools/prof_step.py C2 3 > gpurun_out/r01j_ncu3.log 2>&1
ls -la gpurun_out
