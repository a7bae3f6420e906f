#!/bin/bash
# tensor-map prefetch + ordered kd-in-N issue: tests, phase probe, A/B of the joint step
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_conv3_tc_gpu.py tests/test_models_gpu.py -m gpu -x -q > gpurun_out/r2b_pytest.log 2>&1
tail -3 gpurun_out/r2b_pytest.log | cut -c1-200
timeout 300 python tools/tc_phase_probe.py > gpurun_out/r2b_tc_phase.txt 2>&1; cut -c1-330 gpurun_out/r2b_tc_phase.txt | head -12
for o in 0 1 0 1; do
  VAESEG_KDN_ORDERED=$o timeout 600 python bench.py --mode joint --kernel-table --no-roofline > gpurun_out/r2b_bench_ord$o.json 2> gpurun_out/r2b_bench_ord$o.err
  echo "ordered=$o $(cut -c1-130 gpurun_out/r2b_bench_ord$o.json)"; grep "tc_kdn_ex  " gpurun_out/r2b_bench_ord$o.err | head -1; grep "vs_conv3x3x3_fprop  " gpurun_out/r2b_bench_ord$o.err | head -1
done
VAESEG_KDN_ORDERED=1 timeout 600 python bench.py --mode seg --no-roofline > gpurun_out/r2b_bench_seg_ord1.json 2>/dev/null; cut -c1-130 gpurun_out/r2b_bench_seg_ord1.json
timeout 600 python bench.py --mode seg --no-roofline > gpurun_out/r2b_bench_seg_ord0.json 2>/dev/null; cut -c1-130 gpurun_out/r2b_bench_seg_ord0.json
