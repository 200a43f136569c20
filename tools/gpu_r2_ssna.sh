#!/bin/bash
# round 2: SSNA passes — tests, bench, per-kernel times and instruction counts (ncu launch list)
mkdir -p gpurun_out
O=gpurun_out
echo "== pytest"; timeout 900 python -m pytest tests -x -q -m gpu > $O/pytest_gpu.log 2>&1; echo "rc=$?"; tail -4 $O/pytest_gpu.log
echo "== bench ssna"; timeout 300 python bench.py --steps 20 --warmup 5 --ssna --no-extras > $O/bench_ssna.json 2> $O/bench_ssna.err; echo "rc=$?"
python - <<'P'
import json
j = json.load(open("gpurun_out/bench_ssna.json")); print("ssna %.4f ms" % j["ms_per_step"], "e2e %.4f" % j["e2e"]["ms_per_step"], j.get("parity"))
P
echo "== ncu launch list (ssna)"
timeout 600 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__thread_inst_executed.sum --clock-control none -k regex:"blur_z|ssna_z|shade_pass|render_frame|ssna_post" -s 6 -c 6 --csv --log-file $O/launches_ssna.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extras --ssna > $O/ncu_ssna_list.log 2>&1; echo "rc=$?"
python - <<'P'
import csv
rows = [r for r in csv.reader(open("gpurun_out/launches_ssna.csv")) if len(r) > 10]
h = rows[0]
for r in rows[1:]:
    print(r[h.index("Kernel Name")][:30], r[h.index("Metric Name")], r[h.index("Metric Value")])
P
