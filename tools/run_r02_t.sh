#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -q -m gpu --tb=short 2>&1 | tail -50 > gpurun_out/pytest_gpu_t.log
grep -E "passed|failed|^FAILED" gpurun_out/pytest_gpu_t.log
for sw in 0 1; do
  MDM_SIDE_WGRAD=$sw python bench.py --no-cpu-baseline --steps 8 --warmup 4 --config cc12m_1024x1024 --batch 1 2>/dev/null | python -c "
import sys, json
for line in sys.stdin:
    if line.startswith('{'):
        d=json.loads(line); print('side=$sw', d['value'], d['ms_per_step'], d['config']['workload'])"
done
