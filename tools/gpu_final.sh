#!/bin/bash
# round-end evidence on one GPU: parity tests, bench, ncu launch list + full capture of the shipped kernel
mkdir -p gpurun_out
echo "== pytest gpu"; timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/pytest_gpu.log
echo "== bench"; timeout 900 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; echo "rc=$?"; cat gpurun_out/bench_final.json; tail -2 gpurun_out/bench_final.err
echo "== reference arm"; timeout 900 python bench.py --impl reference --steps 10 --warmup 2 > gpurun_out/bench_reference.json 2>&1; cat gpurun_out/bench_reference.json | cut -c1-300
echo "== ncu launch list"; timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches.csv python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1; echo "rc=$?"
echo "== ncu full"; timeout 600 ncu --set full --clock-control none --import-source on -k regex:render_ -s 4 -c 1 -o gpurun_out/prof_final -f python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_final.log 2>&1; echo "rc=$?"
if [ -n "$1" ]; then
echo "== config 3 depth $1"; timeout 1800 python bench.py --scene iso --depth $1 --width 3840 --height 2160 --steps 20 --warmup 5 > gpurun_out/cfg3_d$1.json 2> gpurun_out/cfg3_d$1.err; echo "rc=$?"; cat gpurun_out/cfg3_d$1.json; tail -3 gpurun_out/cfg3_d$1.err
fi
