#!/bin/bash
# 1-GPU bench over the BASELINE.json configs / env counts; JSON lines -> gpurun_out/matrix_<tag>.jsonl
TAG=${1:-rX}
mkdir -p gpurun_out
OUT=gpurun_out/matrix_$TAG.jsonl
: > $OUT
run() { timeout 400 python bench.py --config $1 --num-envs $2 --steps ${3:-30} --warmup 5 --pool ${4:-4} --no-cpu --no-e2e --no-sweep >> $OUT 2>> gpurun_out/matrix_$TAG.err; }
run command_direction 4096 200
run command_direction 65536 100
run command_direction 1048576 40
run contacts 65536 100
run contacts 1048576 30 3
run rough_terrain 262144 50
run berkeley_humanoid 4096 200
run berkeley_humanoid 65536 100
run berkeley_humanoid 262144 50
run berkeley_humanoid 1048576 30 3
TAG=$TAG python - <<'PY'
import json, os
for line in open(f"gpurun_out/matrix_{os.environ['TAG']}.jsonl"):
    d = json.loads(line)
    r = d["roofline"]
    print(f'{d["config"]["workload"][:60]:60s} {d["ms_per_step"]*1e3:8.1f} us/step {d["value"]/1e9:6.2f} G/s  post {r["kernel_us"]:7.1f} us {r["frac"]*100:5.1f}%  bytes/env {r["bytes_per_env"]}  spec {d["kernel_variant"]["specialised_launches"]}')
PY
