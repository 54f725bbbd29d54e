#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/multicast.log
: > $L
echo "== gemm tests default" >> $L
python -m pytest tests/test_gemm_gpu.py -q -m gpu --tb=line 2>&1 | tail -8 >> $L
echo "== gemm tests one-tile kernel only, cluster 2" >> $L
MDM_GEMM_NO_PERSISTENT=1 timeout 600 python -m pytest tests/test_gemm_gpu.py -q -m gpu --tb=line 2>&1 | tail -8 >> $L
echo "== gemm tests one-tile kernel only, cluster 4" >> $L
MDM_GEMM_NO_PERSISTENT=1 MDM_GEMM_CLUSTER=4 timeout 600 python -m pytest tests/test_gemm_gpu.py -q -m gpu --tb=line 2>&1 | tail -8 >> $L
echo "== net + fullwidth tests (cluster 2 default)" >> $L
timeout 900 python -m pytest tests/test_net_gpu.py tests/test_fullwidth_gpu.py tests/test_attention_gpu.py -q -m gpu --tb=short 2>&1 | tail -15 >> $L
for cs in 1 2 4; do
  export MDM_GEMM_CLUSTER=$cs
  echo "== cluster $cs" >> $L
  python tests/profile_conv.py fwd 64 64 64 256 256 >> $L 2>&1
  python tests/profile_conv.py fwd 64 32 32 512 512 >> $L 2>&1
  python tests/profile_conv.py fwd 64 16 16 768 768 >> $L 2>&1
  python tests/profile_conv.py wgrad 64 64 64 256 256 >> $L 2>&1
  MDM_REPORT_TOP=12 python tests/gemm_shape_report.py cc12m_64x64 64 2>&1 | head -12 >> $L
done
cat $L
