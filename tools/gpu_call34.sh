#!/bin/bash
mkdir -p gpurun_out
for v in 1 0; do
VAESEG_FUSE_APPLY=$v timeout 600 python bench.py --mode seg --kernel-table > gpurun_out/r2h_bench_seg_fa$v.json 2>gpurun_out/r2h_bench_seg_fa$v.err; echo "fuse_apply=$v $(cut -c1-118 gpurun_out/r2h_bench_seg_fa$v.json)"
head -12 gpurun_out/r2h_bench_seg_fa$v.err
grep -E "in_relu|vs_inorm_relu_apply|tc_kdn_ex|vs_conv3x3x3_fprop" gpurun_out/r2h_bench_seg_fa$v.err | sed -n 1,40p
done
