#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_models_gpu.py tests/test_conv3_tc_gpu.py -m gpu -x -q > gpurun_out/r2e_pytest.log 2>&1
tail -5 gpurun_out/r2e_pytest.log | cut -c1-250
grep -E "argmax agreement|bf16 gradients|VAE: argmax" gpurun_out/r2e_pytest.log | head -20
for v in 1 0 1 0; do
VAESEG_INBLOCK_TC=$v timeout 600 python bench.py --mode joint --no-roofline > gpurun_out/r2e_bench_joint_ib$v.json 2>/dev/null; echo "inblock_tc=$v $(cut -c1-118 gpurun_out/r2e_bench_joint_ib$v.json)"
done
for v in 1 0; do
VAESEG_INBLOCK_TC=$v timeout 600 python bench.py --mode vae --no-roofline > gpurun_out/r2e_bench_vae_ib$v.json 2>/dev/null; echo "inblock_tc=$v $(cut -c1-118 gpurun_out/r2e_bench_vae_ib$v.json)"
VAESEG_INBLOCK_TC=$v timeout 600 python bench.py --mode seg --no-roofline > gpurun_out/r2e_bench_seg_ib$v.json 2>/dev/null; echo "inblock_tc=$v $(cut -c1-118 gpurun_out/r2e_bench_seg_ib$v.json)"
done
