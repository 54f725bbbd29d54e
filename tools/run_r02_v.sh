#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -q -m gpu --tb=short 2>&1 | tail -30 > gpurun_out/pytest_gpu_v.log
grep -E "passed|failed|^FAILED" gpurun_out/pytest_gpu_v.log
python bench.py --config cc12m_1024x1024 --batch 1 --steps 8 --warmup 4 --no-cpu-baseline 2>/dev/null | python -c "
import sys, json
for line in sys.stdin:
    if line.startswith('{'):
        d=json.loads(line); print('1024 b1', d['value'], d['ms_per_step'])"
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err
tail -2 gpurun_out/bench_default.err
