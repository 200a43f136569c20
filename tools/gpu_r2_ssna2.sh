#!/bin/bash
# round 2: the fused SSNA kernel (ssna_post) with TMA-prefetched tiles — sanitizers first, then the three forms timed
mkdir -p gpurun_out
O=gpurun_out
echo "== memcheck"; timeout 280 compute-sanitizer --tool memcheck --print-limit 3 python -m pytest tests/test_ssna.py -x -q -m gpu 2>&1 | grep -v "Host Frame\|^=========$" | head -30
echo "== racecheck"; timeout 280 compute-sanitizer --tool racecheck --print-limit 3 python -m pytest tests/test_ssna.py -x -q -m gpu -k "device_render" 2>&1 | grep -E "Race|hazard|at yv|passed|failed|SUMMARY" | head -8
echo "== pytest ssna (plain)"; timeout 300 python -m pytest tests/test_ssna.py tests/test_jitter.py -x -q -m gpu 2>&1 | tail -3
for f in 0 1 2; do
  YV_SSNA_FUSED=$f timeout 300 python bench.py --steps 20 --warmup 5 --ssna --no-extras > $O/bench_ssna_f$f.json 2> $O/bench_ssna_f$f.err; echo "fused=$f rc=$?"
done
python - <<'P'
import json
for f in (0, 1, 2):
    try:
        j = json.load(open("gpurun_out/bench_ssna_f%d.json" % f)); print("fused", f, "ssna %.4f ms" % j["ms_per_step"], "launches", j["gpu_launches"], j.get("parity"))
    except Exception as e: print(f, "ERR", e)
P
