#!/bin/bash
# compute-sanitizer over the tiny-shape kernel tests of the hand-rolled protocols (mbarrier rings, TMEM double buffering,
# the cross-CTA shift hand-off, TMA loads): memcheck + racecheck + synccheck.  Run on a GPU box:
#   gpurun --timeout 1500 -- 'bash tools/sanitize.sh'          logs -> gpurun_out/sanitize_*.log (copy into profiles/)
mkdir -p gpurun_out
SEL='test_tc_fprop_and_dgrad and (case0 or case2 or case6) or test_tc_kdn_fprop_and_dgrad and case0 or test_k2s2_tensor_core_gather_and_scatter and case0 or test_tc_wgrad and case0'
for tool in memcheck racecheck synccheck; do
  timeout 300 compute-sanitizer --tool $tool --error-exitcode 9 --launch-timeout 120 \
      python -m pytest tests/test_conv3_tc_gpu.py -x -q -k "$SEL" > gpurun_out/sanitize_$tool.log 2>&1
  echo "$tool rc=$?" | tee -a gpurun_out/sanitize_$tool.log
  grep -E "ERROR SUMMARY|passed|failed|Error|hazard" gpurun_out/sanitize_$tool.log | tail -5
done
