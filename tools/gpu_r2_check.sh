#!/bin/bash
# round 2, call 1: the whole GPU test suite (new: test_multi_device.py) and the unchanged bench line
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
nvidia-smi topo -m > gpurun_out/topo.txt 2>&1
echo "== pytest multi" ; timeout 600 python -m pytest tests/test_multi_device.py -x -q -m gpu > gpurun_out/pytest_multi.log 2>&1 ; echo "rc=$?" ; tail -25 gpurun_out/pytest_multi.log
echo "== pytest gpu" ; timeout 1200 python -m pytest tests -q -m gpu --deselect tests/test_multi_device.py > gpurun_out/pytest_gpu.log 2>&1 ; echo "pytest rc=$?" ; tail -15 gpurun_out/pytest_gpu.log
echo "== smoke" ; timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1 ; echo "smoke rc=$?" ; tail -3 gpurun_out/smoke.log
echo "== bench" ; timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err ; echo "rc=$?" ; cat gpurun_out/bench_c2.json ; tail -3 gpurun_out/bench_c2.err
