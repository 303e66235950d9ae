"""
Stress of the persistent slab loop with the shortest possible slab iteration (the entity phase alone:
two small TMA loads, two TMA stores, a few hundred ns per slab) -- the launch shape that exposed the
uniform-register clobber of the mbarrier init value (csrc/device_utils.cuh, profiles/r2_01_*).

    GFB_DEBUG=8 GFB_TILE=32 python tools/loop_stress.py [num_envs] [repetitions]

GFB_DEBUG=8 forces the loop for every slab size.  Prints "ok" or the failure.
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("GFB_SPEC_JIT", "0")
import torch  # noqa: E402

import genesis_forge_b200 as gfb  # noqa: E402
from configs import specs  # noqa: E402
from configs.env_builder import build_env, dropin_namespace  # noqa: E402


def main() -> int:
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 200_000
    reps = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    dev = torch.device("cuda", 0)
    gfb.set_device(dev)
    env = build_env(specs.get(os.environ.get("SPEC", "contacts")), dropin_namespace(), n, dev, pool=2, seed=5, n_contacts=8)
    try:
        env.build()
        for _ in range(reps):
            env._fused.cache_entity()
        torch.cuda.synchronize()
        same = torch.equal(env.robot_manager.base_pos, env.robot.get_pos()) and torch.equal(
            env.robot_manager.base_quat, env.robot.get_quat())
        print(f"tile={os.environ.get('GFB_TILE', 'auto')} n={n} reps={reps}: ok, copies equal: {same}", flush=True)
        return 0 if same else 1
    except Exception as e:  # noqa: BLE001
        print(f"tile={os.environ.get('GFB_TILE', 'auto')} n={n}: FAILED {type(e).__name__}: {str(e)[:120]}", flush=True)
        return 1


if __name__ == "__main__":
    sys.exit(main())
