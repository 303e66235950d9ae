"""Wall-clock timeline of the Python side of env.step() (where the GPU waits for the host)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import make_dropin_env
from configs import specs

name = sys.argv[1] if len(sys.argv) > 1 else "command_direction"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 1048576
dev = torch.device("cuda", 0)
env = make_dropin_env(specs.get(name), n, dev, 4, 1)
acts = [torch.randn(n, env.action_space.shape[0], device=dev) for _ in range(4)]
fused = env._fused
T = {}


def wrap(obj, attr, key):
    fn = getattr(obj, attr)

    def timed(*a, **k):
        t0 = time.perf_counter()
        r = fn(*a, **k)
        T[key] = T.get(key, 0.0) + time.perf_counter() - t0
        return r

    setattr(obj, attr, timed)


class TimedLib:
    """ctypes library proxy: time spent inside each C entry point."""

    def __init__(self, lib):
        self._lib = lib
        self._cache = {}

    def __getattr__(self, name):
        fn = self._cache.get(name)
        if fn is None:
            raw = getattr(self._lib, name)

            def fn(*a, _raw=raw, _key=f"    C {name}"):
                t0 = time.perf_counter()
                r = _raw(*a)
                T[_key] = T.get(_key, 0.0) + time.perf_counter() - t0
                return r

            self._cache[name] = fn
        return fn


fused.lib = TimedLib(fused.lib)
wrap(fused, "_engine_buffers", "    _engine_buffers")
wrap(fused, "_set_program", "    _set_program")
wrap(fused, "_obs_buffers", "    _obs_buffers")
wrap(fused, "_maybe_specialise", "    _maybe_specialise")
if env.managers["action"] is not None:
    wrap(env.managers["action"], "reset", "    action.reset")
    wrap(env.managers["action"], "_delayed", "    action._delayed")
wrap(env.robot, "control_dofs_position", "    robot.control_dofs_position")
for em in env.managers["entity"]:
    wrap(em, "reset", "    entity.reset")
wrap(env, "_begin_step", "_begin_step")
wrap(fused, "begin_step", "begin_step")
wrap(fused, "action_step", "action_step (launch)")
wrap(env.scene, "step", "scene.step")
wrap(fused, "post_physics", "post_physics (launch + report wait)")
wrap(env, "_host_reset", "host reset handlers")
wrap(fused, "observe", "observe (launch)")
wrap(fused, "finish_logging", "finish_logging")
wrap(env, "_publish", "publish")
wrap(env, "_step_outputs", "step outputs")
for i in range(20):
    env.step(acts[i % 4])
torch.cuda.synchronize()
T.clear()
PROFILE = os.environ.get('GFB_TIMELINE_PROFILE', '0') == '1'  # (per-kernel events cost ~2 us of host time per launch)
if PROFILE:
    fused.profile(True)
K = 100
t0 = time.perf_counter()
for i in range(K):
    env.step(acts[i % 4])
torch.cuda.synchronize()
total = (time.perf_counter() - t0) / K * 1e6
print(f"{name} n={n}: {total:.1f} us/step wall")
acc = 0.0
for k, v in T.items():
    print(f"  {k:40s} {v / K * 1e6:8.1f} us")
    if not k.startswith("    "):
        acc += v / K * 1e6
print(f"  {'(other python in step)':40s} {total - acc:8.1f} us")
if PROFILE:
    prof = fused.profile_read(); aux = fused.profile_read_aux()
    print("  small kernels:", ", ".join(f"{k} {v['kernel_us']:.1f} us x{v['launches']}" for k, v in aux.items()))
    print(f"  kernels: action {prof['action_ms'] / max(prof['action_launches'], 1) * 1e3:.1f} us, post "
          f"{prof['post_ms'] / max(prof['post_launches'], 1) * 1e3:.1f} us")
