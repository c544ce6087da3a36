#!/bin/bash
# usage: bash scripts/gpu_multi.sh N TAG  -- short multi-GPU lines of both configs (frames: incl. the train-sharded matcher)
N=$1; TAG=$2
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --frames 128 --steps 3 --warmup 3 --noise-frames 32 > gpurun_out/${TAG}_frames_${N}gpu.json 2> gpurun_out/${TAG}_frames_${N}gpu.err
echo "frames rc=$?"; tail -c 1800 gpurun_out/${TAG}_frames_${N}gpu.json; tail -4 gpurun_out/${TAG}_frames_${N}gpu.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --config stream --stream-frames 16 --steps 3 --warmup 3 > gpurun_out/${TAG}_stream_${N}gpu.json 2> gpurun_out/${TAG}_stream_${N}gpu.err
echo "stream rc=$?"; tail -c 1200 gpurun_out/${TAG}_stream_${N}gpu.json; tail -4 gpurun_out/${TAG}_stream_${N}gpu.err
