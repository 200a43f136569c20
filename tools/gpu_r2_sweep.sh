#!/bin/bash
# round 2: the variant libraries in the tree on four workloads (no tests, no bench): tools/variant.sh builds them
mkdir -p gpurun_out
O=gpurun_out
for lib in yoxel-voxel_b200/libyv_b200*.so; do
  lib=$(basename $lib); n=${lib%.so}
  for rep in 1 2; do
  YV_B200_LIB=$lib timeout 300 python tools/sweep.py --persistent 0 --frames 20 --out $O/sw_${n}_c2_$rep.json 2>&1 | grep -o '"ms_median": [0-9.]*\|"same_image": [a-z]*' | tr '\n' ' ' | sed "s/^/$n c2 /"; echo
  done
  YV_B200_LIB=$lib timeout 300 python tools/sweep.py --secondary --persistent 0 --frames 8 --out $O/sw_${n}_c4.json 2>&1 | grep -o '"ms_median": [0-9.]*' | sed "s/^/$n c4 /"
  YV_B200_LIB=$lib timeout 300 python tools/sweep.py --persistent 0 --frames 8 --width 7680 --height 4320 --out $O/sw_${n}_8k.json 2>&1 | grep -o '"ms_median": [0-9.]*' | sed "s/^/$n 8k /"
  YV_B200_LIB=$lib timeout 300 python tools/sweep.py --persistent 0 --frames 8 --scene iso --depth 11 --width 3840 --height 2160 --pos 0.2,0.15,0.45 --dir 0.6,0.7,-0.45 --out $O/sw_${n}_iso.json 2>&1 | grep -o '"ms_median": [0-9.]*' | sed "s/^/$n iso /"
done
