#!/bin/bash
# round 2, GPU call 1: unverified tests, precision on conditioned weights, bench modes
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/c1_smi.txt 2>&1
VAESEG_TEST_UNVERIFIED=1 VAESEG_TEST_KDN=1 timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/c1_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/c1_pytest.log
timeout 600 python tools/precision_cond.py > gpurun_out/c1_precision.log 2>&1
timeout 600 python tools/precision_cond.py --patch 64 --seg-steps 0 --vae-steps 0 --modes seg,vae --precisions bf16 > gpurun_out/c1_precision_k0.log 2>&1
for m in joint seg vae; do
  timeout 600 python bench.py --mode $m --kernel-table > gpurun_out/c1_bench_$m.json 2> gpurun_out/c1_bench_$m.err
done
timeout 600 python bench.py --mode joint --no-e2e-prefetch --no-roofline --no-cpu-baseline > gpurun_out/c1_bench_joint_serial.json 2> gpurun_out/c1_bench_joint_serial.err
timeout 900 python bench.py --mode joint_ttt --no-roofline --steps 10 > gpurun_out/c1_bench_ttt.json 2> gpurun_out/c1_bench_ttt.err
VAESEG_KDN=1 timeout 600 python bench.py --mode joint --no-roofline --no-cpu-baseline > gpurun_out/c1_bench_joint_kdn.json 2> gpurun_out/c1_bench_joint_kdn.err
tail -3 gpurun_out/c1_pytest.log
cat gpurun_out/c1_bench_*.json | cut -c1-400
