#!/bin/bash
# single GPU: persistent-kernel multicast check + ncu --set full captures of the step's kernels
mkdir -p gpurun_out
L=gpurun_out/multicast_persistent.log
: > $L
MDM_GEMM_CLUSTER=2 python -m pytest tests/test_gemm_gpu.py -q -m gpu --tb=line 2>&1 | tail -5 >> $L
for cs in 1 2; do
  MDM_GEMM_CLUSTER=$cs python tests/profile_conv.py fwd 64 64 64 256 256 >> $L 2>&1
  MDM_GEMM_CLUSTER=$cs MDM_REPORT_TOP=3 python tests/gemm_shape_report.py cc12m_64x64 64 2>&1 | head -9 >> $L
done
cat $L
export MDM_NO_GRAPH=1
NCU="ncu --set full --clock-control none --import-source on --profile-from-start off"
run() { name=$1; shift; timeout 300 $NCU "$@" -o gpurun_out/prof_$name python tests/profile_step.py cc12m_64x64 64 train > gpurun_out/ncu_$name.log 2>&1; }
run conv -k regex:gemm_tc_persistent -s 8 -c 4
run conv_long -k regex:"gemm_tc_kernel<0, 0>" -s 0 -c 3
run wgrad -k regex:"gemm_tc_kernel<1, 1>" -s 2 -c 2
run attn_fwd -k regex:attn_fwd_kernel -s 0 -c 2
run attn_bwd -k regex:attn_bwd_kernel -s 0 -c 2
run gn_fwd -k regex:"gn_apply_kernel|gn_stats_kernel" -s 0 -c 4
run gn_bwd -k regex:"gn_bwd_apply_kernel|gn_bwd_reduce_kernel|cast_colsum_kernel" -s 0 -c 6
timeout 300 ncu --set full --clock-control none -k regex:adam_ema_sweep -s 1 -c 1 -o gpurun_out/prof_sweep python tests/profile_optim.py > gpurun_out/ncu_sweep.log 2>&1
ls -la gpurun_out/*.ncu-rep
