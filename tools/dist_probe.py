"""Host-side timing of the sharded step's sections (run under torchrun)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
from bench import make_dropin_env
from configs import specs
import genesis_forge_b200.fused as F
rank = int(os.environ.get("RANK", 0)); lr = int(os.environ.get("LOCAL_RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1))
dev = torch.device("cuda", lr); torch.cuda.set_device(dev)
if world > 1: dist.init_process_group("nccl", device_id=dev)
n = 1 << 20
env = make_dropin_env(specs.get("command_direction"), n, dev, 4, 1234 + rank)
fused = env._fused
if world > 1:
    fused.shard(dist.group.WORLD, n * world)
acts = [torch.randn(n, 12, device=dev) for _ in range(4)]
T = {}
def wrap(obj, name):
    fn = getattr(obj, name)
    def w(*a, **k):
        t0 = time.perf_counter(); r = fn(*a, **k); T[name] = T.get(name, 0.0) + time.perf_counter() - t0; return r
    setattr(obj, name, w)
for nm in ["action_step", "post_physics", "observe", "finish_logging", "_allreduce_logging", "_engine_buffers", "_set_program"]:
    wrap(fused, nm)
wrap(env, "_host_reset"); wrap(env, "_publish")
for i in range(10): env.step(acts[i % 4])
torch.cuda.synchronize(); T.clear()
K = 100
t0 = time.perf_counter()
for i in range(K): env.step(acts[i % 4])
torch.cuda.synchronize()
tot = (time.perf_counter() - t0) / K * 1e6
if rank == 0:
    print(f"world {world}: {tot:.1f} us/step; sections (us/step): " + ", ".join(f"{k} {v / K * 1e6:.1f}" for k, v in sorted(T.items(), key=lambda kv: -kv[1])))
if world > 1: dist.destroy_process_group()
