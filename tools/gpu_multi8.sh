#!/bin/bash
N=${1:-8}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo$N.txt 2>&1
run() { timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N "$@"; }
echo "== frames p2p";  run --steps 50 --warmup 5 > gpurun_out/multi_frames_p2p_$N.json 2> gpurun_out/multi_frames_p2p_$N.err; echo rc=$?; cat gpurun_out/multi_frames_p2p_$N.json | cut -c1-700; grep -v OMP gpurun_out/multi_frames_p2p_$N.err | tail -3
echo "== frames nccl"; run --steps 50 --warmup 5 --gather nccl > gpurun_out/multi_frames_nccl_$N.json 2> gpurun_out/multi_frames_nccl_$N.err; echo rc=$?; cat gpurun_out/multi_frames_nccl_$N.json | cut -c1-300
echo "== tiles 8K fractal p2p"; run --steps 30 --warmup 5 --partition tiles --width 7680 --height 4320 > gpurun_out/multi_tiles_p2p_$N.json 2> gpurun_out/multi_tiles_p2p_$N.err; echo rc=$?; cat gpurun_out/multi_tiles_p2p_$N.json | cut -c1-300
echo "== config 5: iso d13 8K 64-frame flythrough, tiles p2p"; run --scene iso --depth 13 --steps 64 --warmup 3 --partition tiles --flythrough --width 7680 --height 4320 > gpurun_out/cfg5_$N.json 2> gpurun_out/cfg5_$N.err; echo rc=$?; cat gpurun_out/cfg5_$N.json | cut -c1-900; grep -v OMP gpurun_out/cfg5_$N.err | tail -3
