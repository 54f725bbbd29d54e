#!/bin/bash
# CTA-pair (cta_group::2) one-tile kernel: parity cases and the GEMM micro-bench, pairs on / off
mkdir -p gpurun_out
MDM_GEMM_PAIR=1 MDM_GEMM_NO_PERSISTENT=1 timeout 400 python tests/gemm_cases.py --bench > gpurun_out/pair_cases.txt 2>&1
echo "rc=$?" >> gpurun_out/pair_cases.txt
MDM_GEMM_NO_PERSISTENT=1 timeout 300 python tests/gemm_cases.py --bench > gpurun_out/pair_cases_off.txt 2>&1
MDM_GEMM_PAIR=2 MDM_GEMM_NO_PERSISTENT=1 timeout 300 python tests/gemm_cases.py > gpurun_out/pair_cases_2.txt 2>&1
grep -v PASS gpurun_out/pair_cases.txt | tail -30
