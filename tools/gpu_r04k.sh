#!/bin/bash
TAG=r04k
mkdir -p gpurun_out
python - <<PY | tee gpurun_out/${TAG}_selftest.log
import importlib, sys
sys.path.insert(0, '.')
fm = importlib.import_module("bachelor-thesis_b200")
c = fm.Context(64, 64)
print("selftest mismatches:", c.selftest_division(1 << 26, 11))
c.close()
PY
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -4 | tee gpurun_out/${TAG}_tests.log
for c in C1 C2 C3; do timeout 600 python tools/ab_probe.py $c 2>&1 | cut -c1-200 | tee -a gpurun_out/${TAG}_ab.log; done
timeout 600 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none -k regex:k_march_first -c 3 --csv --log-file gpurun_out/${TAG}_first.csv python tools/prof_step.py C2 3 > /dev/null 2>&1
grep "k_march_first" gpurun_out/${TAG}_first.csv | tail -2 | cut -c280-420
