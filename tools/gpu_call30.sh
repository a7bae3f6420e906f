#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_conv3_tc_gpu.py tests/test_models_gpu.py tests/test_ops_gpu.py -m gpu -x -q > gpurun_out/r2d_pytest.log 2>&1
tail -5 gpurun_out/r2d_pytest.log | cut -c1-250
for i in 1 2; do
timeout 600 python bench.py --mode joint --kernel-table > gpurun_out/r2d_bench_joint$i.json 2> gpurun_out/r2d_bench_joint$i.err; cut -c1-118 gpurun_out/r2d_bench_joint$i.json
done
grep -E "kdn_planar|head_conv|vs_conv3x3x3_dgrad " gpurun_out/r2d_bench_joint1.err | head
timeout 600 python bench.py --mode seg --no-roofline > gpurun_out/r2d_bench_seg.json 2>/dev/null; cut -c1-118 gpurun_out/r2d_bench_seg.json
timeout 600 python bench.py --mode vae --no-roofline > gpurun_out/r2d_bench_vae.json 2>/dev/null; cut -c1-118 gpurun_out/r2d_bench_vae.json
