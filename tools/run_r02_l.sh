#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -q -m gpu --tb=short 2>&1 | tail -80 > gpurun_out/pytest_gpu_l.log
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err
MDM_NO_GRAPH=1 timeout 300 ncu --set full --clock-control none --profile-from-start off -k regex:"gn_bwd_apply|gn_bwd_reduce|cast_colsum|gn_apply" -s 0 -c 8 -o gpurun_out/prof_staged python tests/profile_step.py cc12m_64x64 64 train > gpurun_out/ncu_staged.log 2>&1
MDM_NO_GRAPH=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file gpurun_out/launches_64_b64.csv python tests/profile_step.py cc12m_64x64 64 train > gpurun_out/ncu_64.log 2>&1
python tests/summarize_launches.py gpurun_out/launches_64_b64.csv > gpurun_out/launches_64_b64_summary.txt 2>&1
grep -E "passed|failed|^FAILED" gpurun_out/pytest_gpu_l.log; tail -3 gpurun_out/bench_default.err
