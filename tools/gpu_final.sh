#!/bin/bash
# (under gpurun) end-of-round measurement set: DRAM traffic per workload, the default bench line, ncu launch
# list + full captures of the post kernel (configs 2, 3, 5) and of the re-observation kernel.
TAG=${1:-r2z}
mkdir -p gpurun_out
timeout 900 python tools/measure_traffic.py > gpurun_out/${TAG}_traffic.log 2>&1; tail -3 gpurun_out/${TAG}_traffic.log | cut -c1-200
cp gpurun_out/traffic.json profiles/traffic.json
timeout 900 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; tail -c 2500 gpurun_out/${TAG}_bench.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_${TAG}.csv \
    python bench.py --steps 6 --warmup 3 --no-sweep --no-cpu --no-e2e --no-configs > gpurun_out/ncu_launch_${TAG}.log 2>&1
for cfg in command_direction contacts berkeley_humanoid; do
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:post_kernel -s 4 -c 1 -f -o gpurun_out/post_${TAG}_${cfg} \
      python bench.py --steps 4 --warmup 3 --no-sweep --no-cpu --no-e2e --no-configs --config $cfg > gpurun_out/ncu_${TAG}_${cfg}.log 2>&1
  grep -o '"libraries": \[[^]]*\]' gpurun_out/ncu_${TAG}_${cfg}.log
done
timeout 400 ncu --set full --clock-control none --import-source on -k regex:observe_kernel -s 2 -c 1 -f -o gpurun_out/observe_${TAG} \
    python bench.py --steps 4 --warmup 3 --no-sweep --no-cpu --no-e2e --no-configs > gpurun_out/ncu_${TAG}_obs.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:action_kernel -s 4 -c 1 -f -o gpurun_out/action_${TAG} \
    python bench.py --steps 4 --warmup 3 --no-sweep --no-cpu --no-e2e --no-configs > gpurun_out/ncu_${TAG}_act.log 2>&1
ls gpurun_out | grep ${TAG} | tr '\n' ' '
