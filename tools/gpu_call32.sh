#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2f_pytest.log 2>&1
tail -4 gpurun_out/r2f_pytest.log | cut -c1-250
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
for v in 1 0; do
VAESEG_INBLOCK_TC=$v timeout 600 python bench.py --mode joint --no-roofline > gpurun_out/r2f_bench_joint_ib$v.json 2>/dev/null; echo "inblock_tc=$v $(cut -c1-118 gpurun_out/r2f_bench_joint_ib$v.json)"
done
