"""
Build-time tooling: ahead-of-time specialisations of the fused kernel for the shipped term tables
(the BASELINE.json configs + test tables of configs/specs.py), both slab-size classes, Philox and
injected-draw modes.  Runs without a GPU: a host-only handle describes each table's structure
(genesis_forge_b200/spec.py), nvcc cross-compiles for sm_100a.

    python tools/prebuild_specs.py [config ...]

Kept outside the package: the shipped term tables are workload definitions (configs/), not product code.
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def prebuild(spec_names=None, sizes=(4096, 65536), rng_modes=(1, 0), jobs: int = 8, verbose: bool = False):
    import torch

    import genesis_forge_b200 as gfb
    from genesis_forge_b200 import spec
    from configs import specs
    from configs.env_builder import build_env, dropin_namespace

    prev = gfb.gs.device
    gfb.set_device("cpu")
    todo: dict[str, str] = {}
    try:
        names = list(spec_names or list(specs.ALL) + ["second_entity"])
        if "second_entity" in names:  # two EntityManagers (configs/second_entity.py), at its test's size
            from configs import second_entity

            names.remove("second_entity")
            ns = dropin_namespace()
            for stock_terms in (False, True):
                env = second_entity.add_prop(
                    build_env(second_entity.spec(stock_terms), ns, 16, torch.device("cpu"), n_contacts=8), ns)
                env._dry_run = True
                env.build()
                todo.update(spec.collect(env, rng_modes))
        for name in names:
            table = specs.get(name)
            for n in sizes:
                for n_contacts in {8 if table["contacts"] else 0, 8}:
                    env = build_env(table, dropin_namespace(), n, torch.device("cpu"), n_contacts=n_contacts)
                    env._dry_run = True
                    env.build()
                    todo.update(spec.collect(env, rng_modes))
    finally:
        gfb.gs.device = prev
    return spec.compile_many(todo, jobs=jobs, verbose=verbose)


if __name__ == "__main__":
    paths = prebuild(sys.argv[1:] or None)
    print(len(paths), "specialised kernels")
