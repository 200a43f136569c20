#!/bin/bash
# multi-GPU checks on an N-GPU box: bench in frames (weak) and tiles (strong, 8K) partition, p2p and nccl gather
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo.txt 2>&1
run() { timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N "$@"; }
echo "== frames p2p";  run --steps 30 --warmup 5 > gpurun_out/multi_frames_p2p_$N.json 2> gpurun_out/multi_frames_p2p_$N.err; echo rc=$?; cat gpurun_out/multi_frames_p2p_$N.json; tail -3 gpurun_out/multi_frames_p2p_$N.err
echo "== frames nccl"; run --steps 30 --warmup 5 --gather nccl > gpurun_out/multi_frames_nccl_$N.json 2> gpurun_out/multi_frames_nccl_$N.err; echo rc=$?; cat gpurun_out/multi_frames_nccl_$N.json; tail -3 gpurun_out/multi_frames_nccl_$N.err
echo "== tiles 8K p2p"; run --steps 20 --warmup 5 --partition tiles --width 7680 --height 4320 > gpurun_out/multi_tiles_p2p_$N.json 2> gpurun_out/multi_tiles_p2p_$N.err; echo rc=$?; cat gpurun_out/multi_tiles_p2p_$N.json; tail -3 gpurun_out/multi_tiles_p2p_$N.err
echo "== single 8K"; timeout 600 python bench.py --steps 20 --warmup 5 --width 7680 --height 4320 --no-cpu-baseline > gpurun_out/single_8k.json 2> gpurun_out/single_8k.err; echo rc=$?; cat gpurun_out/single_8k.json
