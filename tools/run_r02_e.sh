#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -q -m gpu --tb=short 2>&1 | tail -150 > gpurun_out/pytest_gpu_e.log
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err
MDM_REPORT_TOP=45 python tests/gemm_shape_report.py cc12m_64x64 64 > gpurun_out/gemm_shapes_64.txt 2>&1
MDM_REPORT_TOP=16 python tests/gemm_shape_report.py cc12m_256x256 32 > gpurun_out/gemm_shapes_256.txt 2>&1
MDM_REPORT_TOP=16 python tests/gemm_shape_report.py cc12m_1024x1024 2 > gpurun_out/gemm_shapes_1024.txt 2>&1
MDM_NO_GRAPH=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file gpurun_out/launches_256_b32.csv python tests/profile_step.py cc12m_256x256 32 train > gpurun_out/ncu_256.log 2>&1
python tests/summarize_launches.py gpurun_out/launches_256_b32.csv > gpurun_out/launches_256_b32_summary.txt 2>&1
grep -E "passed|failed|^FAILED" gpurun_out/pytest_gpu_e.log; tail -3 gpurun_out/bench_default.err
