#!/bin/bash
# Round 2: regenerates every measured line under profiles/ on the build that is in the tree (one B200).
#   gpurun --timeout 2400 -- 'bash tools/gpu_r2_all.sh [noprof]'
# then, in the container:  python tools/collect_profiles.py r02
# Every step has its own timeout so a hung kernel cannot eat the lease.
mkdir -p gpurun_out
O=gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > $O/gpu.txt 2>&1
git rev-parse --short=12 HEAD > $O/commit.txt 2>/dev/null
echo "== smoke" ; timeout 300 python __graft_entry__.py smoke > $O/smoke.log 2>&1 ; echo "smoke rc=$?" ; tail -3 $O/smoke.log
echo "== pytest gpu" ; timeout 1800 python -m pytest tests -x -q -m gpu --durations=8 > $O/pytest_gpu.log 2>&1 ; echo "pytest rc=$?" ; tail -14 $O/pytest_gpu.log
echo "== bench N=1 (headline + configs 1, 4, 3)" ; timeout 900 python bench.py --steps 20 --warmup 5 > $O/bench_n1.json 2> $O/bench_n1.err ; echo "rc=$?" ; tail -2 $O/bench_n1.err
echo "== bench reference arm" ; timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > $O/bench_ref.json 2> $O/bench_ref.err ; echo "rc=$?"
echo "== bench ssna" ; timeout 300 python bench.py --steps 20 --warmup 5 --ssna --no-extras > $O/bench_ssna.json 2> $O/bench_ssna.err ; echo "rc=$?"
echo "== bench 8k" ; timeout 300 python bench.py --steps 10 --warmup 3 --width 7680 --height 4320 --no-cpu-baseline > $O/bench_8k.json 2> $O/bench_8k.err ; echo "rc=$?"
python - <<'P'
import json
def show(tag, j):
    r = j["roofline"]
    print("%-10s %8.3f ms %9.1f Mrays/s frac %.3f  e2e %.3f ms  parity %s" % (tag, j["ms_per_step"], j["value"], r["frac"], j["e2e"]["ms_per_step"], j.get("parity")))
for f in ("bench_n1", "bench_ssna", "bench_8k"):
    try:
        j = json.load(open("gpurun_out/%s.json" % f)); show(f, j)
        for k, v in (j.get("configs") or {}).items():
            if "value" in v: show("  " + k, v)
            else: print("  ", k, v)
        if "wall_s" in j: print("   wall_s", j["wall_s"])
    except Exception as e:
        print(f, "ERR", e)
P
if [ "$1" != "noprof" ]; then
echo "== ncu launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file $O/launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extras > $O/ncu_launch_bench.log 2>&1 ; echo "rc=$?"
echo "== ncu full: primary"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:render_frame -s 4 -c 1 -o $O/prof_primary -f python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extras > $O/ncu_primary.log 2>&1 ; echo "rc=$?"
echo "== ncu full: secondary (config 4)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:render_frame -s 4 -c 1 -o $O/prof_sec -f python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extras --secondary > $O/ncu_sec.log 2>&1 ; echo "rc=$?"
echo "== ncu full: ssna passes"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"blur_z|ssna_z|shade_pass" -s 6 -c 6 -o $O/prof_ssna -f python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extras --ssna > $O/ncu_ssna.log 2>&1 ; echo "rc=$?"
fi
ls -la $O | tail -25
