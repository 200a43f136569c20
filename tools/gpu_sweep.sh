#!/bin/bash
mkdir -p gpurun_out
echo "== pytest gpu" ; timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1 ; echo "pytest rc=$?" ; tail -8 gpurun_out/pytest_gpu.log
echo "== sweep" ; timeout 900 python tools/sweep.py "$@" 2>&1 | tail -40
