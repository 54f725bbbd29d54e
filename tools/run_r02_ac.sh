#!/bin/bash
# full GPU test suite + 1024 b1 + default bench on the build with CTA pairs / TMA epilogue operands / new kernel-form rule
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu --tb=short -x 2>&1 | tail -40 > gpurun_out/pytest_gpu_v9.log
grep -E "passed|failed|^FAILED" gpurun_out/pytest_gpu_v9.log
timeout 300 python bench.py --config cc12m_1024x1024 --batch 1 --steps 8 --warmup 4 --no-cpu-baseline 2>/dev/null | python -c "
import sys, json
for line in sys.stdin:
    if line.startswith('{'):
        d=json.loads(line); print('1024 b1', d['value'], d['ms_per_step'])"
timeout 900 python bench.py > gpurun_out/bench_default_v9.json 2> gpurun_out/bench_default_v9.err
tail -2 gpurun_out/bench_default_v9.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_default_v9.json').read().strip().splitlines()[-1])
print('headline', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d['clocks'], 'roofline', d['roofline']['achieved'], d['roofline']['frac'])
for k,v in d.get('configs',{}).items(): print(k, v.get('value'), v.get('ms_per_step'))
PY
