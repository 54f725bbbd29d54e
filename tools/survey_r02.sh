#!/bin/bash
# Round-2 survey: do the 256 / 1024 configs run, how long is a step, and where does the time go.
mkdir -p gpurun_out
for spec in "cc12m_256x256 32" "cc12m_1024x1024 1" "cc12m_1024x1024 2" "cc12m_1024x1024 8" "cc12m_256x256 16 infer" "cc12m_64x64 64"; do
  timeout 300 python tests/quick_bench.py $spec 2>&1 | tail -3
done > gpurun_out/survey_timing.log 2>&1
for spec in "cc12m_256x256 32" "cc12m_1024x1024 2"; do
  set -- $spec
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file gpurun_out/launches_$1_b$2.csv python tests/profile_step.py $1 $2 train > gpurun_out/ncu_$1.log 2>&1
  python tests/summarize_launches.py gpurun_out/launches_$1_b$2.csv > gpurun_out/launches_$1_b$2_summary.txt 2>&1
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file gpurun_out/launches_sample256_b16.csv python tests/profile_step.py cc12m_256x256 16 sample > gpurun_out/ncu_sample.log 2>&1
python tests/summarize_launches.py gpurun_out/launches_sample256_b16.csv > gpurun_out/launches_sample256_b16_summary.txt 2>&1
cat gpurun_out/survey_timing.log
