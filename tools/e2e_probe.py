"""Where does the e2e step (bench.py:time_e2e) spend its time?  CUDA events around the host->device
copies of a step, the kernels and the observation read-back.   python tools/e2e_probe.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from configs import specs

dev = torch.device("cuda", 0)
n = 1 << 20
env = bench.make_dropin_env(specs.get("command_direction"), n, dev, 4, 1234)
acts = [torch.randn(n, env._fused.D).pin_memory() for _ in range(4)]
marks = []
orig = bench.torch.cuda.Event


def ev():
    e = torch.cuda.Event(enable_timing=True)
    e.record(torch.cuda.current_stream(dev))
    return e


scene = env.scene
fused = env._fused
# instrument: wrap scene.step after time_e2e replaces it -> patch action/post launches instead
a0 = fused.action_step
p0 = fused.post_physics


def action_step(actions):
    marks.append(("before action kernel", ev()))
    a0(actions)
    marks.append(("after action kernel", ev()))


def post_physics(*a, **k):
    marks.append(("state landed", ev()))
    r = p0(*a, **k)
    marks.append(("post kernel done (report read)", ev()))
    return r


fused.action_step = action_step
fused.post_physics = post_physics
ms, h2d, d2h = bench.time_e2e(env, acts, 6, 3, False)
torch.cuda.synchronize()
print(f"e2e {ms:.3f} ms/step, h2d {h2d / 1e6:.1f} MB, d2h {d2h / 1e6:.1f} MB")
# last full step: consecutive marks
per = 4
last = marks[-per * 3:]
for (n0, e0), (n1, e1) in zip(last[:-1], last[1:]):
    print(f"  {n0:32s} -> {n1:32s} {e0.elapsed_time(e1):7.3f} ms")
