#!/bin/bash
# The one-GPU evidence set of profiles/ (sanitizer, bench lines of every BASELINE config, ncu launch list, CUPTI timeline):
#   gpurun --timeout 3000 -- "bash tools/measure_all.sh"; results land in gpurun_out/
mkdir -p gpurun_out
bash tools/sanitize.sh
for m in joint seg vae; do
  timeout 600 python bench.py --mode $m --kernel-table > gpurun_out/r2_bench_$m.json 2> gpurun_out/r2_bench_$m.err
  cut -c1-130 gpurun_out/r2_bench_$m.json
done
timeout 600 python bench.py --mode vae --patch 128 --no-roofline > gpurun_out/r2_bench_vae128.json 2> /dev/null; cut -c1-130 gpurun_out/r2_bench_vae128.json
timeout 900 python bench.py --mode joint_ttt --no-roofline --steps 10 > gpurun_out/r2_bench_joint_ttt_128.json 2> /dev/null; cut -c1-130 gpurun_out/r2_bench_joint_ttt_128.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench_reference.json 2> /dev/null; cut -c1-130 gpurun_out/r2_bench_reference.json
timeout 600 ncu --profile-from-start off --clock-control none --metrics gpu__time_duration.sum --csv --log-file gpurun_out/r2_launches_step_ncu.csv python tools/one_step.py > gpurun_out/r2_one_step.log 2>&1
python tools/agg_ncu.py gpurun_out/r2_launches_step_ncu.csv > gpurun_out/r2_launches_step_summary.txt 2>&1; head -12 gpurun_out/r2_launches_step_summary.txt | cut -c1-100
timeout 600 python tools/timeline.py --tag r2final > gpurun_out/r2_timeline_cupti_summary.txt 2>&1; sed -n 3,8p gpurun_out/r2_timeline_cupti_summary.txt
