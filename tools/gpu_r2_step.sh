#!/bin/bash
# round 2: one regression + variant step on the build in the tree (one B200)
#   tests (ssna, pack, parity sizes), bench N=1 without extras, SSNA bench, then the variant libraries on four workloads
mkdir -p gpurun_out
O=gpurun_out
echo "== pytest (subset)"; timeout 900 python -m pytest tests -x -q -m gpu > $O/pytest_gpu.log 2>&1; echo "rc=$?"; tail -4 $O/pytest_gpu.log
echo "== bench n1"; timeout 600 python bench.py --steps 20 --warmup 5 --no-extras > $O/bench_n1_quick.json 2> $O/bench_n1_quick.err; echo "rc=$?"
echo "== bench ssna"; timeout 300 python bench.py --steps 20 --warmup 5 --ssna --no-extras --no-cpu-baseline > $O/bench_ssna.json 2> $O/bench_ssna.err; echo "rc=$?"
python - <<'P'
import json
for f in ("bench_n1_quick", "bench_ssna"):
    try:
        j = json.load(open("gpurun_out/%s.json" % f)); print(f, "%.4f ms" % j["ms_per_step"], "e2e %.4f" % j["e2e"]["ms_per_step"], j.get("parity"))
    except Exception as e: print(f, "ERR", e)
P
for lib in yoxel-voxel_b200/libyv_b200*.so; do
  lib=$(basename $lib); n=${lib%.so}
  for rep in 1 2; do
  YV_B200_LIB=$lib timeout 300 python tools/sweep.py --persistent 0 --frames 20 --out $O/sw_${n}_c2_$rep.json 2>&1 | grep -o '"ms_median": [0-9.]*' | sed "s/^/$n c2 /"
  done
  YV_B200_LIB=$lib timeout 300 python tools/sweep.py --secondary --persistent 0 --frames 8 --out $O/sw_${n}_c4.json 2>&1 | grep -o '"ms_median": [0-9.]*' | sed "s/^/$n c4 /"
  YV_B200_LIB=$lib timeout 300 python tools/sweep.py --persistent 0 --frames 8 --width 7680 --height 4320 --out $O/sw_${n}_8k.json 2>&1 | grep -o '"ms_median": [0-9.]*' | sed "s/^/$n 8k /"
  YV_B200_LIB=$lib timeout 300 python tools/sweep.py --persistent 0 --frames 8 --scene iso --depth 11 --width 3840 --height 2160 --pos 0.2,0.15,0.45 --dir 0.6,0.7,-0.45 --out $O/sw_${n}_iso.json 2>&1 | grep -o '"ms_median": [0-9.]*' | sed "s/^/$n iso /"
done
