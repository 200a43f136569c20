#!/bin/bash
# round 2: octant culling — parity suite, then config 2 / 4 / 8K with culling on and off
mkdir -p gpurun_out
echo "== pytest gpu" ; timeout 1500 python -m pytest tests -q -x -m gpu > gpurun_out/pytest_gpu.log 2>&1 ; echo "pytest rc=$?" ; tail -15 gpurun_out/pytest_gpu.log
for cull in "" "--no-cull"; do
  tag=cull1; [ -n "$cull" ] && tag=cull0
  echo "== config 2 $tag" ; timeout 600 python bench.py --steps 20 --warmup 5 $cull > gpurun_out/bench_c2_$tag.json 2> gpurun_out/bench_c2_$tag.err ; echo "rc=$?" ; python - <<P
import json
j=json.load(open("gpurun_out/bench_c2_$tag.json")); r=j["roofline"]
print(j["ms_per_step"], j["value"], r["frac"], r["node_visits_per_ray"], r["kernel_node_fetches_per_ray"], j["e2e"]["ms_per_step"], j["parity"])
P
  echo "== config 4 $tag" ; timeout 600 python bench.py --steps 10 --warmup 3 --secondary $cull > gpurun_out/bench_c4_$tag.json 2> gpurun_out/bench_c4_$tag.err ; echo "rc=$?" ; python - <<P
import json
j=json.load(open("gpurun_out/bench_c4_$tag.json")); r=j["roofline"]
print(j["ms_per_step"], j["value"], r["frac"], r["node_visits_per_ray"], r["kernel_node_fetches_per_ray"], j["parity"])
P
  echo "== config 2 @8K $tag" ; timeout 600 python bench.py --steps 10 --warmup 3 --width 7680 --height 4320 --no-cpu-baseline $cull > gpurun_out/bench_8k_$tag.json 2> gpurun_out/bench_8k_$tag.err ; echo "rc=$?" ; python - <<P
import json
j=json.load(open("gpurun_out/bench_8k_$tag.json")); r=j["roofline"]
print(j["ms_per_step"], j["value"], r["frac"], r["node_visits_per_ray"], r["kernel_node_fetches_per_ray"])
P
  echo "== iso d11 4K $tag" ; timeout 600 python bench.py --steps 10 --warmup 3 --scene iso --depth 11 --width 3840 --height 2160 --no-cpu-baseline $cull > gpurun_out/bench_iso_$tag.json 2> gpurun_out/bench_iso_$tag.err ; echo "rc=$?" ; python - <<P
import json
j=json.load(open("gpurun_out/bench_iso_$tag.json")); r=j["roofline"]
print(j["ms_per_step"], j["value"], r["frac"], r["node_visits_per_ray"], r["kernel_node_fetches_per_ray"])
P
done
