#!/bin/bash
TAG=${1:-r03i}
mkdir -p gpurun_out
CS=/usr/local/cuda/bin/compute-sanitizer
run() { name=$1; shift; timeout 1200 "$@" > gpurun_out/${TAG}_san_$name.log 2>&1; echo "$name rc=$? $(grep 'SUMMARY' gpurun_out/${TAG}_san_$name.log | tail -1) | $(grep -E 'passed|failed' gpurun_out/${TAG}_san_$name.log | tail -1)"; }
run memcheck_parity $CS --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -x -q -k "not full_size"
run racecheck_aniso $CS --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_aniso.py -x -q -k "golden or bit_exact or march" 
run racecheck_parity $CS --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -x -q -k "golden"
run memcheck_misc $CS --tool memcheck --error-exitcode 9 python -m pytest tests/test_smooth.py tests/test_record.py tests/test_bgeo.py tests/test_gpu_interop.py -x -q -m gpu
