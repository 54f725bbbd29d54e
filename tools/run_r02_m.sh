#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -q -m gpu --tb=short 2>&1 | tail -60 > gpurun_out/pytest_gpu_m.log
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err
MDM_REPORT_TOP=10 python tests/gemm_shape_report.py cc12m_1024x1024 1 > gpurun_out/gemm_shapes_1024_b1.txt 2>&1
MDM_NO_GRAPH=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file gpurun_out/launches_sample.csv python tests/profile_step.py cc12m_256x256 16 sample > gpurun_out/ncu_sample.log 2>&1
python tests/summarize_launches.py gpurun_out/launches_sample.csv > gpurun_out/launches_sample_summary.txt 2>&1
grep -E "passed|failed|^FAILED" gpurun_out/pytest_gpu_m.log; tail -3 gpurun_out/bench_default.err
