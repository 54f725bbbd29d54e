#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_graph_gpu.py tests/test_net_gpu.py tests/test_fullwidth_gpu.py tests/test_diffusion_gpu.py tests/test_gemm_gpu.py -q -m gpu --tb=short 2>&1 | tail -250 > gpurun_out/pytest_gpu_d.log
tail -5 gpurun_out/pytest_gpu_d.log
