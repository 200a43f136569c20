#!/bin/bash
# refresh the single-GPU bench lines (configs 2, 1, 4, 3 at depth 13, SSNA) with the current build; no profiler
mkdir -p gpurun_out
run() { name=$1; shift; echo "== $name: $*"; timeout 1200 python bench.py "$@" > gpurun_out/$name.json 2> gpurun_out/$name.err; echo "rc=$?"; cut -c1-700 gpurun_out/$name.json; tail -2 gpurun_out/$name.err; }
run cfg2 
run cfg1 --depth 10 --width 512 --height 512 --steps 50
run cfg4 --secondary --steps 30 --warmup 5
run ssna --ssna --steps 50 --warmup 5
run cfg3_d13 --scene iso --depth 13 --width 3840 --height 2160 --steps 20 --warmup 5
