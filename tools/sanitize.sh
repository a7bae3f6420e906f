#!/bin/bash
# compute-sanitizer over the tiny-shape kernel tests of the hand-rolled protocols (mbarrier rings, TMEM multi-buffering,
# the cross-CTA shift hand-off, TMA loads, the padded kd-in-N packs): memcheck + racecheck + synccheck.  The cooperative
# conv + InstanceNorm + ReLU launch (in-kernel grid barrier with a poll limit) runs under memcheck only: the other tools
# slow the CTAs down by orders of magnitude, which the barrier's trap-instead-of-hang limit would report as a failure.
# Run on a GPU box:   gpurun --timeout 1800 -- 'bash tools/sanitize.sh'      logs -> gpurun_out/sanitize_*.log
mkdir -p gpurun_out
SEL='test_tc_fprop_and_dgrad and (case0 or case2 or case6) or test_tc_kdn_fprop_and_dgrad and case0 or test_k2s2_tensor_core_gather_and_scatter and case0 or test_tc_wgrad and case0 or test_tc_kdn_planar_head_and_inblock_dgrad and case0 or test_batched_repack_matches_per_layer_packs'
FUSED='test_tc_conv_instance_norm_relu_in_one_launch and (case0 or case4) and True'
for tool in memcheck racecheck synccheck; do
  sel="$SEL"
  if [ $tool = memcheck ]; then sel="$SEL or $FUSED"; fi
  timeout 420 compute-sanitizer --tool $tool --error-exitcode 9 --launch-timeout 120 \
      python -m pytest tests/test_conv3_tc_gpu.py -x -q -k "$sel" > gpurun_out/sanitize_$tool.log 2>&1
  echo "$tool rc=$?" | tee -a gpurun_out/sanitize_$tool.log
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed|Error|hazard" gpurun_out/sanitize_$tool.log | tail -5
done
