#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -q -m gpu --tb=short 2>&1 | tail -60 > gpurun_out/pytest_gpu_o.log
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
MDM_NO_GRAPH=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file gpurun_out/launches_sample.csv python tests/profile_step.py cc12m_256x256 16 sample > gpurun_out/ncu_sample.log 2>&1
python tests/summarize_launches.py gpurun_out/launches_sample.csv > gpurun_out/launches_sample_summary.txt 2>&1
grep -E "passed|failed|^FAILED" gpurun_out/pytest_gpu_o.log; tail -2 gpurun_out/bench_default.err; tail -2 gpurun_out/smoke.log
