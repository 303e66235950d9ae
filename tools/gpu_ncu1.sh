#!/bin/bash
# (under gpurun) one `ncu --set full` capture of post_kernel: args <tag> [bench args]
TAG=${1:-r2x}; shift
mkdir -p gpurun_out
timeout 400 ncu --set full --clock-control none --import-source on -k regex:post_kernel -s 4 -c 1 -f -o gpurun_out/post_${TAG} \
    python bench.py --steps 4 --warmup 3 --no-sweep --no-cpu --no-e2e --no-configs "$@" > gpurun_out/ncu_${TAG}.log 2>&1
grep -o '"libraries": \[[^]]*\]' gpurun_out/ncu_${TAG}.log
