#!/bin/bash
# 2 GPUs: gradient all-reduce after backward vs overlapped with the (segment-graph) backward, with and without SMs
# reserved for NCCL. One JSON line per run in gpurun_out/overlap_n2.log
mkdir -p gpurun_out
L=gpurun_out/overlap_n2.log
: > $L
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
port=29500
one() {  # env..., then bench args
  port=$((port+1))
  echo "### $*" >> $L
  env "$@" > /dev/null 2>&1 || true
}
runb() { desc=$1; shift; port=$((port+1)); echo "### $desc" >> $L; env $ENVV $TR --master-port $port bench.py --gpus 2 --no-cpu-baseline --only headline "$@" 2>> gpurun_out/overlap_n2.err | python -c "
import sys, json
for line in sys.stdin:
    line=line.strip()
    if line.startswith('{'):
        d=json.loads(line); print(json.dumps({k:d[k] for k in ('value','ms_per_step','n_gpus')} | {'e2e':d['e2e']['value'],'workload':d['config']['workload']}))
" >> $L; }
ENVV="X=1" runb "64x64 b64: all-reduce after backward" --steps 8 --warmup 5
ENVV="MDM_OVERLAP=1 MDM_SM_RESERVE=0" runb "64x64 b64: overlap, no SM reserve" --steps 8 --warmup 5
ENVV="MDM_OVERLAP=1 MDM_SM_RESERVE=8" runb "64x64 b64: overlap, 8 SMs for NCCL" --steps 8 --warmup 5
ENVV="MDM_OVERLAP=1 MDM_SM_RESERVE=16" runb "64x64 b64: overlap, 16 SMs for NCCL" --steps 8 --warmup 5
ENVV="X=1" runb "1024 b1: all-reduce after backward" --config cc12m_1024x1024 --batch 1 --steps 8 --warmup 5
ENVV="MDM_OVERLAP=1 MDM_SM_RESERVE=8" runb "1024 b1: overlap, 8 SMs for NCCL" --config cc12m_1024x1024 --batch 1 --steps 8 --warmup 5
ENVV="MDM_OVERLAP=1 MDM_SM_RESERVE=8 MDM_BUCKET_MB=256" runb "1024 b1: overlap, 8 SMs, 256 MB buckets" --config cc12m_1024x1024 --batch 1 --steps 8 --warmup 5
cat $L; tail -5 gpurun_out/overlap_n2.err
