#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/conv_micro.log
: > $L
run() { "$@" >> $L 2>&1; }
# 1024-px level, C=32: as is (3 CTAs/SM one-tile form), persistent, and the W-folded equivalent shape
run python tests/profile_conv.py fwd 2 1024 1024 32 32
MDM_PERSIST_MIN_N=0 run python tests/profile_conv.py fwd 2 1024 1024 32 32
run python tests/profile_conv.py fwd 2 1024 512 64 64
MDM_PERSIST_MIN_N=0 run python tests/profile_conv.py fwd 2 1024 512 64 64
run python tests/profile_conv.py fwd 2 1024 256 128 128
MDM_SMEM_NARROW_KB=44 run python tests/profile_conv.py fwd 2 1024 1024 32 32
MDM_SMEM_NARROW_KB=99 run python tests/profile_conv.py fwd 2 1024 1024 32 32
# 256-px level C=64 at batch 32
run python tests/profile_conv.py fwd 32 256 256 64 64
run python tests/profile_conv.py fwd 32 256 128 128 128
MDM_PERSIST_MIN_N=0 run python tests/profile_conv.py fwd 32 256 256 64 64
# weight gradients
for kf in 1 2 4; do run python tests/profile_conv.py wgrad 2 1024 1024 32 32 20 $kf; done
for kf in 1 2 4; do run python tests/profile_conv.py wgrad 32 256 256 64 64 20 $kf; done
run python tests/profile_conv.py wgrad 2 1024 512 64 64 20 4
cat $L
# ncu --set full: one launch each of the narrow conv (one-tile form and persistent) and the tall wgrad
ncu --set full --clock-control none --import-source on -k regex:gemm_tc -s 3 -c 1 -o gpurun_out/prof_conv32 \
    python tests/profile_conv.py fwd 2 1024 1024 32 32 2 > /dev/null 2>&1
MDM_PERSIST_MIN_N=0 ncu --set full --clock-control none --import-source on -k regex:gemm_tc -s 3 -c 1 -o gpurun_out/prof_conv32_persistent \
    python tests/profile_conv.py fwd 2 1024 1024 32 32 2 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:gemm_tc -s 3 -c 1 -o gpurun_out/prof_wgrad32 \
    python tests/profile_conv.py wgrad 2 1024 1024 32 32 2 4 > /dev/null 2>&1
ls -la gpurun_out/*.ncu-rep
