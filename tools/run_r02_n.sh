#!/bin/bash
# 2 GPUs: the full default bench line exactly as the driver launches it
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --impl reference --gpus 2 --steps 3 --warmup 1 > gpurun_out/bench_n2_ref.json 2>> gpurun_out/bench_n2.err
tail -5 gpurun_out/bench_n2.err | cut -c1-300; head -c 600 gpurun_out/bench_n2.json; echo; head -c 300 gpurun_out/bench_n2_ref.json
