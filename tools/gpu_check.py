"""Run the parity harness on every spec and print the error statistics (use under gpurun)."""
import os
import sys
import time
import traceback

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from oracle.parity import ParityRun

names = sys.argv[1].split(",") if len(sys.argv) > 1 else ["simple", "command_direction", "contacts", "rough_terrain", "berkeley_humanoid", "kitchen_sink"]
N = int(sys.argv[2]) if len(sys.argv) > 2 else 256
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 60
dev = torch.device("cuda", 0)
print(torch.cuda.get_device_name(0), "torch", torch.__version__, flush=True)
failed = 0
for name in names:
    t0 = time.time()
    try:
        run = ParityRun(name, N, dev, seed=1234)
        stats = run.run(steps, nan_step=7)
        print(f"PASS {name}: {stats} ({time.time()-t0:.1f}s)", flush=True)
    except Exception as e:
        failed += 1
        print(f"FAIL {name}: {type(e).__name__}: {e}", flush=True)
        traceback.print_exc(limit=6)
sys.exit(1 if failed else 0)
