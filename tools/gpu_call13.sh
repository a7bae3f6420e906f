#!/bin/bash
# 2 GPUs: NCCL equivalence tests + a 2-rank bench line (plain and overlapped all-reduce)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_ddp_gpu.py -m gpu -x -q -s > gpurun_out/c13_ddp_pytest.log 2>&1
echo "rc=$?" >> gpurun_out/c13_ddp_pytest.log; grep -E "rel-L2|passed|failed|rc=" gpurun_out/c13_ddp_pytest.log | tail
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --no-roofline > gpurun_out/c13_bench_2gpu.json 2> gpurun_out/c13_bench_2gpu.err
cut -c1-200 gpurun_out/c13_bench_2gpu.json
VAESEG_DDP_OVERLAP=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --no-roofline > gpurun_out/c13_bench_2gpu_ov.json 2> gpurun_out/c13_bench_2gpu_ov.err
cut -c1-200 gpurun_out/c13_bench_2gpu_ov.json
