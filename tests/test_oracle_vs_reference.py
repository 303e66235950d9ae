"""
Pins the oracle port against the UNMODIFIED reference (imported from /root/reference under stub
third-party modules): every manager buffer, every step output, every logged value and the global
torch RNG stream, bit for bit, step by step.  Skipped where /root/reference does not exist (GPU box).
"""
import pytest
import torch

from configs import specs
from oracle import compare, ref_harness
from configs.env_builder import make_scene
from oracle.manager_port import PortEnv

pytestmark = pytest.mark.skipif(not ref_harness.reference_available(), reason="/root/reference not present")


def _same(x, y):
    return bool(((x == y) | (x.isnan() & y.isnan())).all()) if x.is_floating_point() else bool(torch.equal(x, y))


def _run(name, num_envs, steps, spec_override=None, nan_step=7, toggles=None):
    """`toggles`: {step: {manager kind: enabled}} flips `enabled` flags on both sides before that step."""
    spec = specs.get(name)
    if spec_override:
        spec.update(spec_override)
    torch.manual_seed(0)
    ref = ref_harness.make_reference_env(spec, num_envs)
    ref.build()
    _, ex_r = ref.reset()
    rng_r = torch.get_rng_state()

    torch.manual_seed(0)
    scene, terrain, robot = make_scene(spec, torch.device("cpu"), copy_on_get=True)
    port = PortEnv(spec, num_envs, scene, terrain, robot)
    port.build()
    _, ex_p = port.reset()
    rng_p = torch.get_rng_state()
    assert torch.equal(rng_r, rng_p), "RNG stream diverged during build/reset"

    def check(where, ex_r, ex_p):
        bad = compare.diff_exact(compare.reference_snapshot(ref), port.snapshot())
        bad += compare.diff_exact(compare.extras_to_cpu(ex_r), compare.extras_to_cpu(ex_p))
        for g in ex_r["observations"]:
            if not _same(ex_r["observations"][g], ex_p["observations"][g]):
                bad.append(f"obs[{g}]")
        assert not bad, f"{where}: {bad}"

    check("reset", ex_r, ex_p)
    gen = torch.Generator().manual_seed(5)
    resets = 0
    for i in range(steps):
        a = torch.randn(num_envs, 12, generator=gen)
        if i == nan_step:
            a[3, 2] = float("nan")
        for kind, enabled in (toggles or {}).get(i, {}).items():
            getattr(ref, f"{kind}_manager").enabled = enabled
            (port.disabled.discard if enabled else port.disabled.add)(kind)
        torch.set_rng_state(rng_r)
        o_r = ref.step(a.clone())
        rng_r = torch.get_rng_state()
        torch.set_rng_state(rng_p)
        o_p = port.step(a.clone())
        rng_p = torch.get_rng_state()
        assert torch.equal(rng_r, rng_p), f"RNG stream diverged at step {i}"
        for j in range(4):
            assert _same(o_r[j], o_p[j]), f"step {i} output {j}"
        assert ("terminations" in o_r[4]) == ("terminations" in o_p[4]), f"step {i} extras keys"
        check(f"step {i}", o_r[4], o_p[4])
        resets += int((o_r[2] | o_r[3]).sum())
    return resets


@pytest.mark.parametrize("name", list(specs.ALL) + list(specs.VARIANTS))
def test_port_is_bit_identical_to_reference(name):
    assert _run(name, 48, 60) > 0


def test_port_matches_reference_through_timeouts():
    # 1 s episodes (50 steps +-10 %): every env times out at least once in 130 steps
    assert _run("command_direction", 24, 130, {"max_episode_length_sec": 1}) >= 48


def test_port_matches_reference_within_limits_action_manager():
    spec = specs.get("command_direction")
    action = dict(spec["action"], type="within_limits")
    for key in ("scale", "use_default_offset"):
        action.pop(key, None)
    assert _run("command_direction", 16, 30, {"action": action}) >= 0


def test_port_matches_reference_with_managers_disabled():
    """`enabled = False` on the action / termination / reward managers: their step() returns early."""
    toggles = {5: {"reward": False}, 12: {"reward": True, "termination": False}, 20: {"termination": True, "action": False},
               27: {"action": True}}
    assert _run("contacts", 32, 40, toggles=toggles, nan_step=None) > 0
