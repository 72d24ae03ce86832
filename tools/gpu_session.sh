#!/bin/bash
# one gpurun session: logs under gpurun_out/ (usage: tools/gpu_session.sh <tag>)
tag=${1:-s}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/${tag}_smi.txt 2>&1
nproc >> gpurun_out/${tag}_smi.txt
