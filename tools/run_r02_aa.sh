#!/bin/bash
# persistent epilogue with TMA-loaded operands: parity cases, FFN micro-bench (on / off), per-shape report
mkdir -p gpurun_out
timeout 300 python tests/gemm_cases.py > gpurun_out/op_cases.txt 2>&1
timeout 200 python tests/profile_ffn.py > gpurun_out/ffn_micro_op.txt 2>&1
MDM_EPI_OP_TMA=0 timeout 200 python tests/profile_ffn.py > gpurun_out/ffn_micro_op_off.txt 2>&1
MDM_REPORT_TOP=30 timeout 300 python tests/gemm_shape_report.py cc12m_64x64 64 > gpurun_out/gemm_shapes_64_op.txt 2>&1
grep -c PASS gpurun_out/op_cases.txt; grep -v PASS gpurun_out/op_cases.txt | head -12
grep "train" gpurun_out/ffn_micro_op.txt | head -12; head -10 gpurun_out/gemm_shapes_64_op.txt
