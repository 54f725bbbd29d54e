#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -q -m gpu --tb=short 2>&1 | tail -60 > gpurun_out/pytest_gpu_s.log
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
grep -E "passed|failed|^FAILED" gpurun_out/pytest_gpu_s.log; tail -2 gpurun_out/bench_default.err; tail -1 gpurun_out/smoke.log
