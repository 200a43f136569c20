#!/bin/bash
# BASELINE configs 3 and 4 on one GPU, plus schedule comparisons for the incoherent secondary rays
mkdir -p gpurun_out
nproc > gpurun_out/host.txt; free -g >> gpurun_out/host.txt
echo "== config 4: secondary, tiles";      timeout 900 python bench.py --secondary --steps 30 --warmup 5 > gpurun_out/cfg4_tiles.json 2> gpurun_out/cfg4_tiles.err; echo rc=$?; cut -c1-900 gpurun_out/cfg4_tiles.json; tail -2 gpurun_out/cfg4_tiles.err
echo "== config 4: secondary, persistent"; timeout 900 python bench.py --secondary --steps 30 --warmup 5 --schedule persistent --no-cpu-baseline > gpurun_out/cfg4_pers.json 2> gpurun_out/cfg4_pers.err; echo rc=$?; cut -c1-400 gpurun_out/cfg4_pers.json; tail -2 gpurun_out/cfg4_pers.err
echo "== config 3: iso depth ${1:-13} 4K"; timeout 1500 python bench.py --scene iso --depth ${1:-13} --width 3840 --height 2160 --steps 20 --warmup 5 > gpurun_out/cfg3.json 2> gpurun_out/cfg3.err; echo rc=$?; cut -c1-1800 gpurun_out/cfg3.json; tail -2 gpurun_out/cfg3.err
echo "== config 1: depth 10 512x512"; timeout 600 python bench.py --depth 10 --width 512 --height 512 --steps 50 > gpurun_out/cfg1.json 2> gpurun_out/cfg1.err; echo rc=$?; cut -c1-600 gpurun_out/cfg1.json
