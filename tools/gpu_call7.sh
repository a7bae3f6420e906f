#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q -s > gpurun_out/c7_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/c7_pytest.log
grep -E "bf16 gradients|argmax agreement|rel-L2|passed|failed|rc=" gpurun_out/c7_pytest.log | tail -40
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/c7_smoke.log 2>&1; tail -2 gpurun_out/c7_smoke.log
timeout 600 python bench.py --mode joint --kernel-table > gpurun_out/c7_bench_joint.json 2> gpurun_out/c7_bench_joint.err
cut -c1-330 gpurun_out/c7_bench_joint.json
