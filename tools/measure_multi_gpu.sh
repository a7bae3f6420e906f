#!/bin/bash
# Weak scaling at 8 and 4 GPUs (joint step and the joint+TTT config):  gpurun --gpus 8 --timeout 1500 -- "bash tools/measure_multi_gpu.sh"
mkdir -p gpurun_out
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8"
timeout 600 $T --no-roofline > gpurun_out/r2_bench_joint_8gpu.json 2> gpurun_out/r2_bench_joint_8gpu.err; cut -c1-130 gpurun_out/r2_bench_joint_8gpu.json
timeout 600 $T --no-roofline --mode joint_ttt --steps 10 > gpurun_out/r2_bench_joint_ttt_128_8gpu.json 2> gpurun_out/r2_bench_joint_ttt_128_8gpu.err; cut -c1-130 gpurun_out/r2_bench_joint_ttt_128_8gpu.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 4 --no-roofline > gpurun_out/r2_bench_joint_4gpu.json 2> gpurun_out/r2_bench_joint_4gpu.err; cut -c1-130 gpurun_out/r2_bench_joint_4gpu.json
