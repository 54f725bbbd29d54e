#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -q -m gpu --tb=short 2>&1 | tail -60 > gpurun_out/pytest_gpu_k.log
python tests/net_cases.py > gpurun_out/net_parity_tiny.log 2>&1
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err
MDM_GN_LEGACY=1 python bench.py --steps 10 --warmup 3 --only headline --no-cpu-baseline > gpurun_out/bench_headline_gn_legacy.json 2>> gpurun_out/bench_default.err
MDM_NO_GRAPH=1 timeout 300 ncu --set full --clock-control none --profile-from-start off -k regex:"gn_bwd_apply|gn_bwd_reduce" -s 0 -c 4 -o gpurun_out/prof_gn_bwd_staged python tests/profile_step.py cc12m_64x64 64 train > gpurun_out/ncu_gn_bwd_staged.log 2>&1
grep -E "passed|failed|^FAILED" gpurun_out/pytest_gpu_k.log; tail -3 gpurun_out/bench_default.err
