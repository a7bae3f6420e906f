#!/bin/bash
# round 2, GPU call 2: k2s2 tcgen05 kernels (parity + microbench), KDN default, precision on conditioned weights
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_conv3_tc_gpu.py -x -q -k "k2s2" > gpurun_out/c2_k2test.log 2>&1
echo "rc=$?" >> gpurun_out/c2_k2test.log
tail -5 gpurun_out/c2_k2test.log
timeout 300 python tools/k2bench.py 96 > gpurun_out/c2_k2bench.txt 2>&1
cat gpurun_out/c2_k2bench.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/c2_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/c2_pytest.log
tail -4 gpurun_out/c2_pytest.log
timeout 600 python tools/precision_cond.py --precisions bf16 > gpurun_out/c2_precision64.log 2>&1
timeout 900 python tools/precision_cond.py --precisions bf16 --patch 96 --batch 2 --modes seg,joint > gpurun_out/c2_precision96.log 2>&1
timeout 600 python bench.py --mode joint --kernel-table > gpurun_out/c2_bench_joint.json 2> gpurun_out/c2_bench_joint.err
VAESEG_K2_TC=0 timeout 600 python bench.py --mode joint --no-roofline --no-cpu-baseline > gpurun_out/c2_bench_joint_nok2tc.json 2> gpurun_out/c2_bench_joint_nok2tc.err
cat gpurun_out/c2_bench_joint*.json | cut -c1-300
