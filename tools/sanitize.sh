#!/bin/bash
# compute-sanitizer passes over a short parity run of every kernel (under gpurun):
#   bash tools/sanitize.sh [memcheck|racecheck|synccheck|initcheck]
TOOL=${1:-memcheck}
mkdir -p gpurun_out
cat > /tmp/gfb_sanitize_run.py <<'PY'
import sys, os
sys.path.insert(0, os.getcwd())
import torch
from oracle.parity import ParityRun
for name, n, env in (("contacts", 96, {}), ("rough_terrain", 64, {}), ("berkeley_humanoid", 70, {"GFB_NO_SPEC": "1"})):
    os.environ.pop("GFB_NO_SPEC", None)
    os.environ.update(env)
    run = ParityRun(name, num_envs=n, device=torch.device("cuda", 0), seed=11)
    stats = run.run(steps=6, nan_step=2)
    print(name, "steps", stats["steps"], "resets", stats["resets"], run.env._fused.spec_stats())
PY
timeout 1200 compute-sanitizer --tool $TOOL --error-exitcode 9 --print-limit 20 python /tmp/gfb_sanitize_run.py > gpurun_out/sanitize_$TOOL.log 2>&1
echo "exit $?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|steps" gpurun_out/sanitize_$TOOL.log | tail -8
