#!/bin/bash
# 8 GPUs: headline config (plain / overlapped all-reduce) and BASELINE config[3] (128^3, dynamic lambda + TTT)
mkdir -p gpurun_out
run() { tag=$1; shift; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $PORT bench.py --gpus 8 "$@" > gpurun_out/c14_$tag.json 2> gpurun_out/c14_$tag.err; cut -c1-220 gpurun_out/c14_$tag.json; PORT=$((PORT+1)); }
PORT=29600
run joint_8gpu --no-roofline
VAESEG_DDP_OVERLAP=1 run joint_8gpu_overlap --no-roofline
run ttt_8gpu --mode joint_ttt --no-roofline --steps 10
