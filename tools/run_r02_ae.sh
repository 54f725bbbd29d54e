#!/bin/bash
# 2 GPUs: headline under torchrun on the final build (CTA-pair kernels + NCCL all-reduce)
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 \
  bench.py --gpus 2 --only headline --no-cpu-baseline > gpurun_out/bench_n2_v9.json 2> gpurun_out/bench_n2_v9.err
tail -3 gpurun_out/bench_n2_v9.err; cut -c1-700 gpurun_out/bench_n2_v9.json
