#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -q -m gpu 2>&1 | tail -40 > gpurun_out/pytest_gpu.log
export MDM_REPORT_TOP=14
for cfg in "cc12m_1024x1024 2" "cc12m_256x256 32"; do
  set -- $cfg
  for v in "default" "MDM_PERSIST_MIN_N=96" "MDM_PERSIST_MIN_N=96 MDM_SMEM_NARROW_KB=64" "MDM_PERSIST_MIN_N=96 MDM_SMEM_NARROW_KB=44"; do
    echo "=== $1 B=$2  [$v]"
    if [ "$v" = "default" ]; then python tests/gemm_shape_report.py $1 $2 2>&1 | tail -30
    else env $v python tests/gemm_shape_report.py $1 $2 2>&1 | tail -30; fi
  done
done > gpurun_out/narrow_knobs.log 2>&1
cat gpurun_out/pytest_gpu.log
