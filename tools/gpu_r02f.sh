#!/bin/bash
TAG=r02f
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -15 > gpurun_out/${TAG}_tests.log
cat gpurun_out/${TAG}_tests.log
timeout 600 python tools/ab_probe.py C2 2>&1 | cut -c1-420
timeout 600 python tools/ab_probe.py C3 2>&1 | cut -c1-420
timeout 600 python tools/ab_probe.py C1 2>&1 | cut -c1-420
