#!/bin/bash
# (under gpurun) one `ncu --set full` capture of post_kernel per config + the launch list of config 2
mkdir -p gpurun_out
TAG=${1:-r2r}
timeout 400 ncu --set full --clock-control none --import-source on -k regex:post_kernel -s 4 -c 1 -f -o gpurun_out/post_${TAG}_cd \
    python bench.py --steps 4 --warmup 3 --no-sweep --no-cpu --no-e2e --no-configs > gpurun_out/ncu_${TAG}_cd.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:post_kernel -s 4 -c 1 -f -o gpurun_out/post_${TAG}_hum \
    python bench.py --steps 4 --warmup 3 --no-sweep --no-cpu --no-e2e --no-configs --config berkeley_humanoid > gpurun_out/ncu_${TAG}_hum.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:observe_kernel -s 2 -c 1 -f -o gpurun_out/observe_${TAG} \
    python bench.py --steps 4 --warmup 3 --no-sweep --no-cpu --no-e2e --no-configs > gpurun_out/ncu_${TAG}_obs.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_${TAG}.csv \
    python bench.py --steps 6 --warmup 3 --no-sweep --no-cpu --no-e2e --no-configs > gpurun_out/ncu_launch_${TAG}.log 2>&1
ls -la gpurun_out/*${TAG}*
