#!/bin/bash
# Usage (under gpurun): bash tools/gpu_profile.sh <tag> [bench args...]
# Runs the GPU parity tests, the bench, an ncu launch list and one full ncu capture of post_kernel.
TAG=${1:-rX}; shift
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 600 python bench.py "$@" > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
tail -3 gpurun_out/bench_$TAG.err; cat gpurun_out/bench_$TAG.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$TAG.csv \
    python bench.py --steps 6 --warmup 3 --no-sweep --no-cpu --no-e2e "$@" > gpurun_out/ncu_launch_$TAG.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:post_kernel -s 4 -c 1 -f -o gpurun_out/post_$TAG \
    python bench.py --steps 4 --warmup 3 --no-sweep --no-cpu --no-e2e "$@" > gpurun_out/ncu_full_$TAG.log 2>&1
tail -1 gpurun_out/ncu_full_$TAG.log | cut -c1-200
