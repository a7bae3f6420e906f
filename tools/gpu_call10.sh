#!/bin/bash
mkdir -p gpurun_out
for pdl in 0 24 64 100 147; do
VAESEG_PDL=$pdl timeout 600 python bench.py --mode joint --no-roofline --no-cpu-baseline --steps 30 > gpurun_out/c10_bench_joint_pdl$pdl.json 2> gpurun_out/c10_bench_joint_pdl$pdl.err
echo "pdl=$pdl $(cut -c1-130 gpurun_out/c10_bench_joint_pdl$pdl.json)"
done
