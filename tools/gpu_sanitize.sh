#!/bin/bash
# compute-sanitizer over the r02 code: memcheck on the build / parity / aniso / sequence tests (small scenes), racecheck +
# synccheck + initcheck on the smoke frame and the side-stream pre-pass test
TAG=${1:-r03h}
mkdir -p gpurun_out
CS=/usr/local/cuda/bin/compute-sanitizer
run() { name=$1; shift; timeout 900 "$@" > gpurun_out/${TAG}_san_$name.log 2>&1; echo "$name rc=$? $(grep -c 'ERROR SUMMARY' gpurun_out/${TAG}_san_$name.log) $(grep 'ERROR SUMMARY' gpurun_out/${TAG}_san_$name.log | tail -3 | tr '\n' ' ') $(tail -1 gpurun_out/${TAG}_san_$name.log | cut -c1-120)"; }
run memcheck_build $CS --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_build.py -x -q
run memcheck_smoke $CS --tool memcheck --error-exitcode 9 python -c "import __graft_entry__ as g; g.smoke()"
run racecheck_smoke $CS --tool racecheck --error-exitcode 9 python -c "import __graft_entry__ as g; g.smoke()"
run synccheck_smoke $CS --tool synccheck --error-exitcode 9 python -c "import __graft_entry__ as g; g.smoke()"
run initcheck_smoke $CS --tool initcheck --error-exitcode 9 python -c "import __graft_entry__ as g; g.smoke()"
run memcheck_aniso $CS --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_aniso.py -x -q -k "not c2 and not C2"
run memcheck_seq $CS --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_sequence.py -x -q
run racecheck_overlap $CS --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_build.py -x -q -k "prepass_beside"
