#!/bin/bash
# profiles of the final build: launch list of one headline training step, ncu --set full of the paired conv kernel and
# of the persistent kernel's TMA-operand epilogues
mkdir -p gpurun_out
MDM_NO_GRAPH=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file gpurun_out/launches_64_b64_v9.csv python tests/profile_step.py cc12m_64x64 64 train > gpurun_out/ncu_64.log 2>&1
python tests/summarize_launches.py gpurun_out/launches_64_b64_v9.csv > gpurun_out/launches_64_b64_v9_summary.txt 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -s 3 -c 2 -f -o gpurun_out/prof_pair_conv \
    python tests/profile_conv.py fwd 64 16 16 768 768 3 > gpurun_out/ncu_pair.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_persistent -s 6 -c 4 -f -o gpurun_out/prof_op_epi \
    python tests/profile_ffn.py --ncu > gpurun_out/ncu_op.log 2>&1
python tests/ncu_summary.py gpurun_out/prof_pair_conv.ncu-rep gpurun_out/prof_op_epi.ncu-rep > gpurun_out/ncu_v9_summary.txt 2>&1
head -20 gpurun_out/launches_64_b64_v9_summary.txt; cat gpurun_out/ncu_v9_summary.txt; tail -3 gpurun_out/ncu_pair.log gpurun_out/ncu_op.log
