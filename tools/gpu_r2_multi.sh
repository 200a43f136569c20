#!/bin/bash
# round 2: the single-handle multi-GPU renderer and the N>1 bench line on a box with >= 2 GPUs
#   usage: tools/gpu_r2_multi.sh N [strong-depth]
N=${1:-2}; SD=${2:-0}     # strong-depth 0 = what the driver gets: auto (14 if the host has the cores and memory)
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo_n$N.txt 2>&1
nvidia-smi nvlink -gt d -i 0 > gpurun_out/nvlink_gt_n$N.txt 2>&1; nvidia-smi nvlink -s -i 0 >> gpurun_out/nvlink_gt_n$N.txt 2>&1
echo "== pytest multi-device" ; timeout 900 python -m pytest tests/test_multi_device.py tests/test_multi_gpu.py -q -x -m gpu > gpurun_out/pytest_multi_n$N.log 2>&1 ; echo "rc=$?" ; tail -5 gpurun_out/pytest_multi_n$N.log
echo "== bench N=$N (torchrun)" ; timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N --steps 20 --warmup 5 --strong-depth $SD > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err ; echo "rc=$?"
tail -c 600 gpurun_out/bench_n$N.err
python - <<P
import json
try:
    j=json.load(open("gpurun_out/bench_n$N.json"))
    print("value", j["value"], "ms", j["ms_per_step"], "e2e", j["e2e"]["value"], "frac", j["roofline"]["frac"])
    print("batch_1gpu", j["batch_1gpu"]); print("parity", j["parity"])
    s=j["strong_8k"]; print("strong", {k:s.get(k) for k in ("error","skipped","depth","scene_build_s","upload_pack_s","ms_per_frame","value","speedup","efficiency","identical_to_1gpu","identical_to_oracle","host_frames_identical_to_device_frames","replicate","imbalance","limit","nvlink")})
    print("strong e2e", s.get("e2e")); print("one_gpu", s.get("one_gpu")); print("wall", j.get("wall_s"))
except Exception as e: print("ERR", e)
P
echo "== host ingest ceiling"; timeout 120 python tools/host_ingest.py
if [ "$N" == "2" ]; then
echo "== ncu: NVLink bytes of the fused peer stores (small strong-only run, 2 GPUs, one process)"
ncu --query-metrics 2>/dev/null | grep -i "nvl" | head -40 > gpurun_out/ncu_nvlink_metrics.txt
timeout 600 ncu --devices 1 --metrics nvltx__bytes.sum,nvlrx__bytes.sum,gpu__time_duration.sum --clock-control none -k regex:render_frame -c 6 --csv --log-file gpurun_out/ncu_nvlink_n2.csv \
    python bench.py --strong-only --gpus 2 --strong-depth 9 --strong-frames 2 --strong-width 1920 --strong-height 1088 > gpurun_out/ncu_nvlink_n2.log 2>&1 ; echo "rc=$?"
head -5 gpurun_out/ncu_nvlink_metrics.txt; grep -c render_frame gpurun_out/ncu_nvlink_n2.csv
fi
