for t in memcheck synccheck racecheck; do echo "-- $t (standard run)"; bash tools/sanitize.sh $t 2>&1 | tail -5; done
cat > /tmp/gfb_sanitize_cd.py <<'PY'
import sys, os
sys.path.insert(0, os.getcwd())
import torch
from oracle.parity import ParityRun
run = ParityRun("command_direction", num_envs=int(os.environ.get("N", "131072")), device=torch.device("cuda", 0), seed=11)
stats = run.run(steps=2)
print("command_direction looped split-groups", "steps", stats["steps"], "resets", stats["resets"], run.env._fused.spec_stats())
PY
for t in memcheck synccheck racecheck; do echo "-- $t (split groups, persistent loop, 131072 envs, specialised 128-env slabs)"; timeout 900 compute-sanitizer --tool $t --error-exitcode 9 --print-limit 20 python /tmp/gfb_sanitize_cd.py 2>&1 | grep -E "ERROR SUMMARY|RACECHECK SUMMARY|looped" | tail -3; done
