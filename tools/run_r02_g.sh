#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -q -m gpu --tb=short 2>&1 | tail -120 > gpurun_out/pytest_gpu_g.log
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err
python bench.py --config cc12m_1024x1024 --batch 1 --steps 5 --no-cpu-baseline > gpurun_out/bench_1024_b1.json 2>> gpurun_out/bench_default.err
MDM_NO_WFOLD=1 python bench.py --config cc12m_1024x1024 --batch 1 --steps 5 --no-cpu-baseline > gpurun_out/bench_1024_b1_nofold.json 2>> gpurun_out/bench_default.err
MDM_REPORT_TOP=16 python tests/gemm_shape_report.py cc12m_1024x1024 2 > gpurun_out/gemm_shapes_1024.txt 2>&1
grep -E "passed|failed|^FAILED" gpurun_out/pytest_gpu_g.log; tail -3 gpurun_out/bench_default.err
