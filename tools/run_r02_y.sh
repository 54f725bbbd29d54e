#!/bin/bash
# CTA pairs: parity of every GEMM case with pairs on (with and without the persistent form), then the knob sweep
mkdir -p gpurun_out
MDM_GEMM_PAIR=1 MDM_GEMM_NO_PERSISTENT=1 timeout 300 python tests/gemm_cases.py > gpurun_out/pair_cases.txt 2>&1
MDM_GEMM_PAIR=2 timeout 300 python tests/gemm_cases.py > gpurun_out/pair_cases_2.txt 2>&1
rm -f gpurun_out/pair_sweep.txt
for PAIR in 0 1; do for KB in 99 130 200; do
  MDM_GEMM_PAIR=$PAIR MDM_SMEM_BUDGET_KB=$KB MDM_GEMM_NO_PERSISTENT=1 timeout 200 python tests/profile_pair.py >> gpurun_out/pair_sweep.txt 2>&1
done; done
grep -c PASS gpurun_out/pair_cases.txt gpurun_out/pair_cases_2.txt; grep -v PASS gpurun_out/pair_cases.txt gpurun_out/pair_cases_2.txt | head; tail -30 gpurun_out/pair_sweep.txt
