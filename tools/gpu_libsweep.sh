#!/bin/bash
# sweep ablation builds (csrc/Makefile `variant`) on config 2 (and optionally config 4)
mkdir -p gpurun_out
for lib in yoxel-voxel_b200/libyv_b200*.so; do
  lib=$(basename $lib)
  echo "== $lib primary"
  YV_B200_LIB=$lib timeout 300 python tools/sweep.py --persistent 0 --frames 10 --out gpurun_out/sweep_${lib%.so}_primary.json 2>&1 | grep -v "^$" | cut -c1-220
  if [ "$1" == "sec" ]; then
  echo "== $lib secondary"
  YV_B200_LIB=$lib timeout 300 python tools/sweep.py --secondary --persistent 0 --frames 5 --out gpurun_out/sweep_${lib%.so}_secondary.json 2>&1 | grep -v "^$" | cut -c1-220
  fi
done
