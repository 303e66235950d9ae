import cProfile, pstats, sys, os, time
sys.path.insert(0, os.getcwd())
import torch
from bench import make_dropin_env
from configs import specs
n = 65536
dev = torch.device("cuda", 0)
env = make_dropin_env(specs.get("gait_trainer"), n, dev, 4, 1)
acts = [torch.randn(n, 12, device=dev) for _ in range(4)]
for i in range(20): env.step(acts[i % 4])
torch.cuda.synchronize()
t0 = time.perf_counter()
for i in range(50): env.step(acts[i % 4])
torch.cuda.synchronize()
print("us/step", (time.perf_counter() - t0) / 50 * 1e6)
pr = cProfile.Profile(); pr.enable()
for i in range(50): env.step(acts[i % 4])
torch.cuda.synchronize()
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(45)
