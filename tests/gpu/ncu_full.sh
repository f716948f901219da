#!/bin/bash
# GPU-box helper: one `ncu --set full` capture of the named kernels on a 2000-read pass (lock-step, one worker).
# usage: bash tests/gpu/ncu_full.sh <tag> <kernel-regex> [launch-count] [skip]
mkdir -p gpurun_out
TAG=$1; RX=$2; CNT=${3:-6}; SKIP=${4:-0}
VM_QUICK_WORKERS=1 timeout 1200 ncu --set full --clock-control none --import-source on -k regex:$RX -s $SKIP -c $CNT \
    -o gpurun_out/prof_$TAG -f python tests/quick_gpu.py 2000 1 > gpurun_out/ncu_$TAG.log 2>&1
tail -3 gpurun_out/ncu_$TAG.log
