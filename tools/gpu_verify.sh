#!/bin/bash
# one GPU: smoke, the GPU parity suite, the bench line (config 2) and the reference arm; no profiler
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
ls oracle/_ref > gpurun_out/ref_libs.txt 2>&1
echo "== smoke"; timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/smoke.log
echo "== pytest gpu"; timeout 1200 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?"; tail -5 gpurun_out/pytest_gpu.log
echo "== bench"; timeout 900 python bench.py > gpurun_out/bench_verify.json 2> gpurun_out/bench_verify.err; echo "rc=$?"; cat gpurun_out/bench_verify.json; tail -2 gpurun_out/bench_verify.err
echo "== reference arm"; timeout 900 python bench.py --impl reference --steps 10 --warmup 2 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; grep -v "^Loading" gpurun_out/bench_reference.json | cut -c1-1500
