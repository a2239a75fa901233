#!/bin/bash
TAG=r03t
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -4 | tee gpurun_out/${TAG}_tests.log
CS=/usr/local/cuda/bin/compute-sanitizer
timeout 600 $CS --tool memcheck --error-exitcode 9 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | grep -E "smoke ok|SUMMARY|Error|error" | cut -c1-150 | tee -a gpurun_out/${TAG}_tests.log
timeout 600 $CS --tool racecheck --error-exitcode 9 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | grep -E "smoke ok|SUMMARY|Error|error" | cut -c1-150 | tee -a gpurun_out/${TAG}_tests.log
timeout 600 $CS --tool synccheck --error-exitcode 9 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | grep -E "smoke ok|SUMMARY|Error|error" | cut -c1-150 | tee -a gpurun_out/${TAG}_tests.log
for c in C1 C2 C3; do timeout 600 python tools/ab_probe.py $c 2>&1 | cut -c1-330 | tee -a gpurun_out/${TAG}_ab.log; done
for v in a_head j_tma; do for c in C2; do
FLUIDMARCH_LIB=$PWD/build_variants/$v/libfluidmarch.so FLUIDMARCH_AB=1 timeout 300 python tools/latency_probe.py $c 40 2>&1 | tail -1 | tee -a gpurun_out/${TAG}_ab.log
done; done
