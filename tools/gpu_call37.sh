#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_ddp_gpu.py -m gpu -x -q -s > gpurun_out/r2_ddp_2gpu_pytest.log 2>&1; tail -3 gpurun_out/r2_ddp_2gpu_pytest.log | cut -c1-200
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --no-roofline > gpurun_out/r2_bench_joint_2gpu.json 2> gpurun_out/r2_bench_joint_2gpu.err; cut -c1-130 gpurun_out/r2_bench_joint_2gpu.json
