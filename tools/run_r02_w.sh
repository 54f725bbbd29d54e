#!/bin/bash
# FFN / Linear GEMM micro-benchmarks: persistent vs one-tile kernel, epilogue variants, warm and cold L2
mkdir -p gpurun_out
python tests/profile_ffn.py > gpurun_out/ffn_micro.txt 2>&1
MDM_GEMM_NO_PERSISTENT=1 python tests/profile_ffn.py > gpurun_out/ffn_micro_onetile.txt 2>&1
tail -5 gpurun_out/ffn_micro.txt
