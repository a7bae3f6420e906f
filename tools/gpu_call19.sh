#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_conv3_tc_gpu.py -x -q 2>&1 | tail -6
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/c19_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/c19_pytest.log; tail -4 gpurun_out/c19_pytest.log | cut -c1-300
timeout 600 python bench.py --mode joint --kernel-table > gpurun_out/c19_bench_joint.json 2> gpurun_out/c19_bench_joint.err
cut -c1-200 gpurun_out/c19_bench_joint.json; head -12 gpurun_out/c19_bench_joint.err
