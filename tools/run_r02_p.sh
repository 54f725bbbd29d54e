#!/bin/bash
# 8 GPUs: the default bench line exactly as the driver launches it (all configs; cc12m_1024x1024 at 1 sample per GPU)
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/bench_n8.json 2> gpurun_out/bench_n8.err
tail -3 gpurun_out/bench_n8.err | cut -c1-300; grep -c '"metric"' gpurun_out/bench_n8.json
