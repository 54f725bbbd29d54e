#!/bin/bash
# CTA pairs on by default: GEMM parity cases, per-shape report of the headline step, default bench
mkdir -p gpurun_out
timeout 300 python tests/gemm_cases.py > gpurun_out/pair_cases_default.txt 2>&1
MDM_REPORT_TOP=45 timeout 300 python tests/gemm_shape_report.py cc12m_64x64 64 > gpurun_out/gemm_shapes_64_pair.txt 2>&1
timeout 900 python bench.py > gpurun_out/bench_pair.json 2> gpurun_out/bench_pair.err
grep -c PASS gpurun_out/pair_cases_default.txt; grep -v PASS gpurun_out/pair_cases_default.txt | head -8
head -12 gpurun_out/gemm_shapes_64_pair.txt; cat gpurun_out/bench_pair.json | cut -c1-1500
