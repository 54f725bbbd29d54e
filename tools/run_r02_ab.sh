#!/bin/bash
# which launches go to the persistent form now that the one-tile form has CTA pairs: sweep of the k-block threshold
mkdir -p gpurun_out
rm -f gpurun_out/persist_kb_sweep.txt
for KB in 48 35 23 11; do
  echo "== MDM_PERSIST_MAX_KBLOCKS=$KB cc12m_64x64 b64" >> gpurun_out/persist_kb_sweep.txt
  MDM_PERSIST_MAX_KBLOCKS=$KB MDM_REPORT_TOP=14 timeout 300 python tests/gemm_shape_report.py cc12m_64x64 64 >> gpurun_out/persist_kb_sweep.txt 2>&1
done
for KB in 48 23; do
  echo "== MDM_PERSIST_MAX_KBLOCKS=$KB cc12m_256x256 b32" >> gpurun_out/persist_kb_sweep.txt
  MDM_PERSIST_MAX_KBLOCKS=$KB MDM_REPORT_TOP=14 timeout 300 python tests/gemm_shape_report.py cc12m_256x256 32 >> gpurun_out/persist_kb_sweep.txt 2>&1
done
grep "==\|^total" gpurun_out/persist_kb_sweep.txt
