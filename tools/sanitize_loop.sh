#!/bin/bash
# compute-sanitizer over the PERSISTENT slab loop (under gpurun): the entity-only launch with the loop forced
# for 32-env slabs (the launch shape of profiles/r2_01) and a full looped step of the contacts config.
#   bash tools/sanitize_loop.sh [memcheck|racecheck|synccheck]
TOOL=${1:-memcheck}
mkdir -p gpurun_out
cat > /tmp/gfb_sanitize_loop.py <<'PY'
import sys, os
sys.path.insert(0, os.getcwd())
import torch
from oracle.parity import ParityRun
run = ParityRun("contacts", num_envs=int(os.environ.get("N", "120000")), device=torch.device("cuda", 0), seed=11)
stats = run.run(steps=2)
print("contacts looped", "steps", stats["steps"], "resets", stats["resets"], run.env._fused.spec_stats())
PY
{
GFB_DEBUG=8 GFB_NO_SPEC=1 GFB_TILE=32 timeout 900 compute-sanitizer --tool $TOOL --error-exitcode 9 --print-limit 20 python tools/loop_stress.py 200000 2 2>&1 | grep -E "ERROR SUMMARY|RACECHECK SUMMARY|tile=" | tail -3
GFB_TILE=64 GFB_NO_SPEC=1 timeout 1500 compute-sanitizer --tool $TOOL --error-exitcode 9 --print-limit 20 python /tmp/gfb_sanitize_loop.py 2>&1 | grep -E "ERROR SUMMARY|RACECHECK SUMMARY|looped" | tail -3
} | tee gpurun_out/sanitize_loop_$TOOL.log
