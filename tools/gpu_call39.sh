#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_models_gpu.py -m gpu -x -q -k "validation or cli_shim or checkpoint" > gpurun_out/r2j_pytest.log 2>&1
tail -15 gpurun_out/r2j_pytest.log | cut -c1-250
for v in 1 0; do
VAESEG_VAL_GRAPH=$v timeout 900 python bench.py --mode joint_ttt --no-roofline --steps 10 > gpurun_out/r2j_bench_ttt_g$v.json 2> gpurun_out/r2j_bench_ttt_g$v.err
python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/r2j_bench_ttt_g$v.json') if l.startswith('{')][-1])
print("val_graph=$v", round(d['value'],1), d['config']['ttt'])
PY
done
