#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_conv3_tc_gpu.py -x -q 2>&1 | tail -4
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/c11_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/c11_pytest.log; tail -3 gpurun_out/c11_pytest.log
for ks in 1 0; do
VAESEG_CONV3_KSPLIT=$ks timeout 600 python bench.py --mode joint --no-roofline --no-cpu-baseline --steps 30 > gpurun_out/c11_bench_ks$ks.json 2> gpurun_out/c11_bench_ks$ks.err
echo "ksplit=$ks $(cut -c1-130 gpurun_out/c11_bench_ks$ks.json)"
done
for ws in 2 3; do
VAESEG_WGRAD_STREAMS=$ws timeout 600 python bench.py --mode joint --no-roofline --no-cpu-baseline --steps 30 > gpurun_out/c11_bench_ws$ws.json 2> gpurun_out/c11_bench_ws$ws.err
echo "wgrad_streams=$ws $(cut -c1-130 gpurun_out/c11_bench_ws$ws.json)"
done
