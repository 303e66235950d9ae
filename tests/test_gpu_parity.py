"""
GPU parity tests proper: the CUDA drop-in, called through the C ABI, against the oracle port on the
same seeded inputs, step by step with resets, resamples, a NaN-action step and (short-episode
variants) timeouts.  Bar: masks / counters / reset indices bit-exact; fp32 within 1e-5 rel + 1e-6.
"""
import pytest
import torch

pytestmark = pytest.mark.gpu

CONFIGS = ["simple", "command_direction", "contacts", "rough_terrain", "berkeley_humanoid", "kitchen_sink"]


@pytest.mark.parametrize("name", CONFIGS)
def test_step_parity(name, cuda_device):
    from oracle.parity import ParityRun

    run = ParityRun(name, num_envs=256, device=cuda_device, seed=1234)
    stats = run.run(steps=120, nan_step=7)
    assert stats["steps"] == 120
    assert stats["resets"] > 0
    print(name, stats)


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
