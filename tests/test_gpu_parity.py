"""
GPU parity tests proper: the CUDA drop-in, called through the C ABI, against the oracle port on the
same seeded inputs, step by step with resets, resamples, a NaN-action step and (short-episode
variants) timeouts.  Bar: masks / counters / reset indices bit-exact; fp32 within 1e-5 rel + 1e-6.
"""
import pytest
import torch

pytestmark = pytest.mark.gpu

CONFIGS = ["simple", "command_direction", "contacts", "rough_terrain", "berkeley_humanoid", "kitchen_sink",
           "custom_terms"]
SPLIT = {"custom_terms"}  # user-defined Python terms: kernel phases run as separate (generic) launches


@pytest.mark.parametrize("name", CONFIGS)
def test_step_parity(name, cuda_device):
    """Default path: the specialised kernel for the config's table structure."""
    from oracle.parity import ParityRun

    run = ParityRun(name, num_envs=256, device=cuda_device, seed=1234)
    stats = run.run(steps=120, nan_step=7)
    assert stats["steps"] == 120
    assert stats["resets"] > 0
    spec_stats = run.env._fused.spec_stats()
    if name in SPLIT:
        assert run.env._fused.split_mode and spec_stats["generic_launches"] == 2 + 4 * 120, spec_stats
    else:
        # (the two generic launches are the build-time entity phase and the initial reset phase)
        assert spec_stats["specialised_launches"] == 120 and spec_stats["generic_launches"] == 2, spec_stats
    print(name, stats, spec_stats)


@pytest.mark.parametrize("name", CONFIGS)
def test_step_parity_generic_interpreter(name, cuda_device, monkeypatch):
    """Same run through the generic (interpreting) kernel: what runs when no specialisation exists."""
    from oracle.parity import ParityRun

    monkeypatch.setenv("GFB_NO_SPEC", "1")
    run = ParityRun(name, num_envs=256, device=cuda_device, seed=1234)
    stats = run.run(steps=60, nan_step=7)
    assert stats["resets"] > 0
    spec_stats = run.env._fused.spec_stats()
    assert spec_stats["specialised_launches"] == 0
    assert spec_stats["generic_launches"] == (2 + 4 * 60 if name in SPLIT else 62)


def test_live_mutation_keeps_the_specialised_kernel(cuda_device):
    """Weights / params / ranges are live values: mutating them must not need a new specialisation."""
    from oracle.parity import ParityRun

    run = ParityRun("command_direction", num_envs=64, device=cuda_device, seed=3)
    run.reset()
    for i in range(10):
        run.step()
    for env_like in (run.env,):
        env_like.reward_manager.cfg["lin_vel_z"].weight = -4.0
        env_like.reward_manager.cfg["base_height_target"].params["target_height"] = 0.33
        env_like.velocity_command.range = {"lin_vel_x": [0.0, 2.0], "lin_vel_y": [-0.5, 0.5], "ang_vel_z": [-1.0, 1.0]}
    run.port.spec["rewards"]["lin_vel_z"]["weight"] = -4.0
    run.port.spec["rewards"]["base_height_target"]["params"]["target_height"] = 0.33
    run.port.command["velocity_command"]["range"] = {"lin_vel_x": [0.0, 2.0], "lin_vel_y": [-0.5, 0.5], "ang_vel_z": [-1.0, 1.0]}
    for i in range(20):
        run.step()
    stats = run.env._fused.spec_stats()
    assert stats["generic_launches"] == 2 and stats["specialised_launches"] == 30 and len(stats["libraries"]) == 1, stats


@pytest.mark.parametrize("name", ["command_direction", "berkeley_humanoid"])
def test_short_episodes_exercise_timeouts(name, cuda_device):
    from oracle.parity import ParityRun

    run = ParityRun(name, num_envs=192, device=cuda_device, seed=99, spec_override={"max_episode_length_sec": 1})
    stats = run.run(steps=150)
    assert stats["resets"] > 192  # every env timed out at least once


@pytest.mark.parametrize("num_envs", [1, 31, 33, 130])
def test_ragged_sizes(num_envs, cuda_device):
    """Slab tails: env counts that are not multiples of the slab size or of 4 (non-TMA path)."""
    from oracle.parity import ParityRun

    run = ParityRun("contacts", num_envs=num_envs, device=cuda_device, seed=5)
    run.run(steps=40)


@pytest.mark.parametrize("name", CONFIGS)
def test_cuda_path_reproduces_reference_golden_trace(name, cuda_device):
    """
    The CUDA path against the committed traces of the UNMODIFIED reference (tests/golden/, generated
    in the build container by oracle/make_golden.py): same seeded inputs, the reference's random draws
    injected via the oracle port (which the CPU suite pins bit-exactly to the same traces).
    """
    import os

    from oracle.make_golden import GOLDEN_DIR
    from oracle.parity import ParityRun, _close

    gold = torch.load(os.path.join(GOLDEN_DIR, f"{name}.pt"), weights_only=False)
    run = ParityRun(name, num_envs=gold["num_envs"], device=cuda_device, seed=gold["seed"], sanitize=False)
    run.reset()
    for i, g in enumerate(gold["step"]):
        out_e, out_p = run.step(nan_action=(i == gold["nan_step"]))
        assert torch.equal(out_e[2].cpu(), g["terminated"]), f"step {i} terminated"
        assert torch.equal(out_e[3].cpu(), g["truncated"]), f"step {i} truncated"
        n = int(g["reset_idx"].numel())
        assert torch.equal(run.env._fused.reset_idx[:n].cpu(), g["reset_idx"]), f"step {i} reset_idx"
        ok, abs_err, _ = _close(out_e[1], g["rewards"])
        assert ok, f"step {i} rewards err {abs_err}"
        for group, want in g["obs"].items():
            ok, abs_err, _ = _close(out_e[4]["observations"][group], want)
            assert ok, f"step {i} obs[{group}] err {abs_err}"
        for key, want in g["logging"].items():
            ok, abs_err, _ = _close(torch.as_tensor(out_e[4]["episode"][key]).reshape(()), want)
            assert ok, f"step {i} extras[{key}] err {abs_err}"


def test_million_env_properties(cuda_device):
    """
    Full-size (1,048,576 envs) run checked through size-independent properties: reset indices are
    exactly the ascending positions of (terminated | truncated); the logged termination fractions
    equal the mask means; episode length is zero exactly on reset envs; observations are finite and
    the command columns of the observation row equal the command buffer.
    """
    import genesis_forge_b200 as gfb
    from oracle import specs
    from oracle.env_builder import build_env, dropin_namespace

    gfb.set_device(cuda_device)
    n = 1 << 20
    env = build_env(specs.get("command_direction"), dropin_namespace(), n, cuda_device, pool=2, seed=7, n_contacts=0,
                    apply_setters=False)
    env.build()
    env.reset()
    for step in range(3):
        actions = torch.randn(n, 12, device=cuda_device)
        obs, rew, term, trunc, extras = env.step(actions)
        done = term | trunc
        idx = done.nonzero().reshape(-1)
        n_reset = env._fused.report.n_reset
        assert n_reset == idx.numel() > 0
        assert torch.equal(env._fused.reset_idx[:n_reset], idx)
        assert torch.equal(env.episode_length == 0, done)
        assert bool(torch.isfinite(obs).all()) and bool(torch.isfinite(rew).all())
        assert torch.equal(obs[:, :3], env.velocity_command.command)
        assert torch.equal(obs[:, 36:48], env.action_manager.get_actions())
        frac = extras["episode"]["Terminations / fall_over"]
        assert abs(float(frac) - float(term.float().mean())) < 1e-6
        assert torch.equal(env.actions[idx], torch.zeros(n_reset, 12, device=cuda_device))
