#!/bin/bash
mkdir -p gpurun_out
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err
python bench.py --config cc12m_1024x1024 --batch 1 --steps 8 --warmup 4 --no-cpu-baseline > gpurun_out/bench_1024_b1.json 2>> gpurun_out/bench_default.err
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
MDM_NO_GRAPH=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file gpurun_out/launches_256_b32.csv python tests/profile_step.py cc12m_256x256 32 train > gpurun_out/ncu_256.log 2>&1
python tests/summarize_launches.py gpurun_out/launches_256_b32.csv > gpurun_out/launches_256_b32_summary.txt 2>&1
MDM_NO_GRAPH=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file gpurun_out/launches_1024_b1.csv python tests/profile_step.py cc12m_1024x1024 1 train > gpurun_out/ncu_1024.log 2>&1
python tests/summarize_launches.py gpurun_out/launches_1024_b1.csv > gpurun_out/launches_1024_b1_summary.txt 2>&1
tail -2 gpurun_out/bench_default.err; tail -1 gpurun_out/smoke.log
