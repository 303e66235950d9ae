"""cProfile of the Python side of env.step() at a small env count (host-bound regime)."""
import cProfile, pstats, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import make_dropin_env
from configs import specs
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
dev = torch.device("cuda", 0)
env = make_dropin_env(specs.get("command_direction"), n, dev, 4, 1)
acts = [torch.randn(n, 12, device=dev) for _ in range(4)]
for i in range(50): env.step(acts[i % 4])
torch.cuda.synchronize()
import time
t0 = time.perf_counter()
for i in range(500): env.step(acts[i % 4])
torch.cuda.synchronize()
print("us/step", (time.perf_counter() - t0) / 500 * 1e6)
pr = cProfile.Profile(); pr.enable()
for i in range(500): env.step(acts[i % 4])
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(28)
