#!/bin/bash
mkdir -p gpurun_out
VAESEG_PDL=1 timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/c9_pytest_pdl.log 2>&1
echo "pytest(pdl) rc=$?" >> gpurun_out/c9_pytest_pdl.log
tail -4 gpurun_out/c9_pytest_pdl.log
timeout 1200 python -m pytest tests -m gpu -x -q -s > gpurun_out/c9_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/c9_pytest.log
grep -E "bf16 gradients|passed|failed|rc=" gpurun_out/c9_pytest.log | tail -12
for pdl in 0 1; do
VAESEG_PDL=$pdl timeout 600 python bench.py --mode joint --no-roofline --no-cpu-baseline > gpurun_out/c9_bench_joint_pdl$pdl.json 2> gpurun_out/c9_bench_joint_pdl$pdl.err
cut -c1-200 gpurun_out/c9_bench_joint_pdl$pdl.json
done
VAESEG_PDL=1 timeout 600 python bench.py --mode seg --no-roofline --no-cpu-baseline > gpurun_out/c9_bench_seg_pdl1.json 2> gpurun_out/c9_bench_seg_pdl1.err
cut -c1-200 gpurun_out/c9_bench_seg_pdl1.json
