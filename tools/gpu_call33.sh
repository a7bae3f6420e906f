#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_conv3_tc_gpu.py -m gpu -x -q -k "in_one_launch" > gpurun_out/r2g_pytest_unit.log 2>&1
tail -5 gpurun_out/r2g_pytest_unit.log | cut -c1-250
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2g_pytest.log 2>&1
tail -4 gpurun_out/r2g_pytest.log | cut -c1-250
for v in 1 0 1 0; do
VAESEG_FUSE_APPLY=$v timeout 600 python bench.py --mode joint --no-roofline > gpurun_out/r2g_bench_joint_fa$v.json 2>gpurun_out/r2g_bench_joint_fa$v.err; echo "fuse_apply=$v $(cut -c1-118 gpurun_out/r2g_bench_joint_fa$v.json)"
done
for m in seg vae; do for v in 1 0; do
VAESEG_FUSE_APPLY=$v timeout 600 python bench.py --mode $m --no-roofline > gpurun_out/r2g_bench_${m}_fa$v.json 2>/dev/null; echo "fuse_apply=$v $(cut -c1-118 gpurun_out/r2g_bench_${m}_fa$v.json)"
done; done
