#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu 2>&1 | tail -25 > gpurun_out/pytest_gpu.log
python tests/net_cases.py > gpurun_out/net_parity_tiny.log 2>&1
python tests/fullwidth_cases.py > gpurun_out/parity_fullwidth.log 2>&1
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err
for g in 0 1; do
  if [ $g = 1 ]; then export MDM_NO_GRAPH=1; else unset MDM_NO_GRAPH; fi
  python bench.py --config cc12m_1024x1024 --batch 1 --steps 5 --no-cpu-baseline > gpurun_out/bench_1024_b1_nograph$g.json 2>> gpurun_out/bench_default.err
done
export MDM_NO_GRAPH=1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file gpurun_out/launches_1024_b1.csv python tests/profile_step.py cc12m_1024x1024 1 train > gpurun_out/ncu_1024_b1.log 2>&1
python tests/summarize_launches.py gpurun_out/launches_1024_b1.csv > gpurun_out/launches_1024_b1_summary.txt 2>&1
cat gpurun_out/pytest_gpu.log; tail -5 gpurun_out/bench_default.err; head -c 1500 gpurun_out/bench_default.json
