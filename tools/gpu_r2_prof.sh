#!/bin/bash
# ncu full captures of render_frame with and without octant culling (config 2)
mkdir -p gpurun_out
for tag in cull1 cull0; do
  flag=""; [ $tag = cull0 ] && flag="--no-cull"
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:render_frame -s 12 -c 1 -o gpurun_out/prof_$tag -f \
     python bench.py --steps 3 --warmup 3 --no-cpu-baseline $flag > gpurun_out/ncu_$tag.log 2>&1 ; echo "$tag rc=$?"
done
ls -la gpurun_out/*.ncu-rep
