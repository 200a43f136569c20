#!/bin/bash
# Run on the GPU box through gpurun: parity tests, smoke, bench (both schedules), ncu launch list + one full capture.
# Every step is wrapped in its own timeout so a hung kernel cannot eat the lease.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
echo "== smoke" ; timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1 ; echo "smoke rc=$?" ; tail -3 gpurun_out/smoke.log
echo "== pytest gpu" ; timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1 ; echo "pytest rc=$?" ; tail -15 gpurun_out/pytest_gpu.log
echo "== bench tiles" ; timeout 600 python bench.py --schedule tiles > gpurun_out/bench_tiles.json 2> gpurun_out/bench_tiles.err ; echo "rc=$?" ; cat gpurun_out/bench_tiles.json ; tail -3 gpurun_out/bench_tiles.err
echo "== bench persistent" ; timeout 600 python bench.py --schedule persistent --no-cpu-baseline > gpurun_out/bench_persistent.json 2> gpurun_out/bench_persistent.err ; echo "rc=$?" ; cat gpurun_out/bench_persistent.json ; tail -3 gpurun_out/bench_persistent.err
if [ "$1" != "noprof" ]; then
echo "== ncu launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch_bench.log 2>&1 ; echo "rc=$?"
echo "== ncu full"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:render_ -s 4 -c 2 -o gpurun_out/prof_tiles -f python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_bench.log 2>&1 ; echo "rc=$?"
fi
ls -la gpurun_out
