#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -q -m gpu --tb=short 2>&1 | tail -60 > gpurun_out/pytest_gpu_q.log
L=gpurun_out/side_wgrad.log
: > $L
one() { desc=$1; shift; echo "### $desc" >> $L; env "$@" 2>> gpurun_out/side_wgrad.err | python -c "
import sys, json
for line in sys.stdin:
    line=line.strip()
    if line.startswith('{'):
        d=json.loads(line); print(json.dumps({'value':d['value'],'ms_per_step':d['ms_per_step'],'workload':d['config']['workload']}))
" >> $L; }
B="python bench.py --no-cpu-baseline --steps 8 --warmup 4"
one "1024 b1 side=0" MDM_SIDE_WGRAD=0 $B --config cc12m_1024x1024 --batch 1
one "1024 b1 side=1" MDM_SIDE_WGRAD=1 $B --config cc12m_1024x1024 --batch 1
one "1024 b2 side=0" MDM_SIDE_WGRAD=0 $B --config cc12m_1024x1024 --batch 2
one "1024 b2 side=1" MDM_SIDE_WGRAD=1 $B --config cc12m_1024x1024 --batch 2
one "1024 b4 side=0" MDM_SIDE_WGRAD=0 $B --config cc12m_1024x1024 --batch 4
one "1024 b4 side=1" MDM_SIDE_WGRAD=1 $B --config cc12m_1024x1024 --batch 4
one "64x64 b64 side=0" MDM_SIDE_WGRAD=0 $B --only headline
one "64x64 b64 side=1" MDM_SIDE_WGRAD=1 $B --only headline
one "256 b32 side=1" MDM_SIDE_WGRAD=1 $B --config cc12m_256x256 --batch 32
one "256 b32 side=0" MDM_SIDE_WGRAD=0 $B --config cc12m_256x256 --batch 32
cat $L; grep -E "passed|failed|^FAILED" gpurun_out/pytest_gpu_q.log; tail -3 gpurun_out/side_wgrad.err | cut -c1-200
