"""
Host-side logic of the drop-in, on CPU: manager API surface, term compiler (config -> packed
program), live config mutation, observation tracing, threshold derivation, error behaviour.
The fused step itself needs a GPU and is covered by the -m gpu tests.
"""
import ctypes
import math

import numpy as np
import pytest
import torch

import genesis_forge_b200 as gfb
from genesis_forge_b200 import _native as nat
from genesis_forge_b200.fused import UnsupportedTermError, asin_tilt_threshold, combine_logging
from genesis_forge_b200.managers import CommandManager, ObservationManager, RewardManager, VelocityCommandManager
from genesis_forge_b200.mdp import rewards
from genesis_forge_b200.rng import ReplayRng
from configs import specs
from configs.env_builder import build_env, dropin_namespace


@pytest.fixture()
def cpu_device():
    prev = gfb.gs.device
    gfb.set_device("cpu")
    yield torch.device("cpu")
    gfb.gs.device = prev


def dry_env(name, n=32, override=None):
    spec = specs.get(name)
    if override:
        spec.update(override)
    env = build_env(spec, dropin_namespace(), n, torch.device("cpu"))
    env._dry_run = True
    env.build()
    env._fused._set_program()
    return env


@pytest.mark.parametrize("name", list(specs.ALL))
def test_every_config_compiles(name, cpu_device):
    env = dry_env(name)
    P = env._fused.program.head
    spec = specs.get(name)
    assert P.num_envs == 32 and P.num_dofs == 12
    assert P.n_reward == len(spec["rewards"])
    assert P.n_termination == len(spec["terminations"])
    assert P.n_command == len(spec["commands"])
    assert P.n_contact == len(spec["contacts"])
    assert P.n_obs_groups == len(spec["observations"])
    # reward weights are fp32(weight * dt) in table order (reward_manager.py:184)
    for r, item in enumerate(spec["rewards"].values()):
        assert P.reward[r].weight == np.float32(item["weight"] * spec["dt"])
    assert env.observation_space.shape[0] == P.obs_group[0].n_cols * P.obs_group[0].history


def test_command_direction_program_contents(cpu_device):
    env = dry_env("command_direction")
    P, K = env._fused.program.head, nat.K
    ops = [P.reward[r].op for r in range(P.n_reward)]
    assert ops == [K["GFB_R_BASE_HEIGHT"], K["GFB_R_TRACK_LIN_VEL"], K["GFB_R_TRACK_ANG_VEL"],
                   K["GFB_R_LIN_VEL_Z"], K["GFB_R_ACTION_RATE"], K["GFB_R_DOF_SIMILAR"]]
    assert P.reward[0].p[0] == np.float32(0.3)
    assert P.reward[1].p[0] == np.float32(0.25) and P.reward[1].mgr == 0
    assert P.termination[0].op == K["GFB_T_TIMEOUT"] and P.termination[0].time_out == 1
    assert P.termination[1].op == K["GFB_T_BAD_ORIENTATION"] and P.termination[1].time_out == 0
    assert P.base_max_episode_length == 1000 and P.max_len_random_span == np.float32(100.0)
    assert P.command[0].resample_steps == 250 and P.command[0].n_dims == 3
    srcs = [env._fused.program.obs_cols[c].src for c in range(48)]
    assert srcs[:3] == [K["GFB_O_COMMAND"]] * 3 and srcs[3:6] == [K["GFB_O_ANG_VEL_B"]] * 3
    assert srcs[12:24] == [K["GFB_O_DOF_POS"]] * 12 and srcs[36:] == [K["GFB_O_TARGETS"]] * 12
    assert env._fused.program.obs_cols[24].scale == np.float32(0.05)
    # action parameters: Go2 joint limits as clip, default pose as offset
    assert P.action_mode == 1 and P.action_scale[0] == np.float32(0.25)
    assert P.action_offset[1] == np.float32(0.8) and P.action_clip_hi[2] == np.float32(-0.83776)


def test_live_config_mutation_is_picked_up(cpu_device):
    env = dry_env("command_direction")
    fused, P = env._fused, env._fused.program.head
    env.reward_manager.cfg["lin_vel_z"].weight = -3.0
    fused._set_program()
    assert P.reward[3].weight == np.float32(-3.0 * env.dt)
    env.reward_manager.cfg["base_height_target"].params["target_height"] = 0.35
    fused._set_program()
    assert P.reward[0].p[0] == np.float32(0.35)
    env.velocity_command.range = {"lin_vel_x": [0.0, 2.0], "lin_vel_y": [-1.0, 1.0], "ang_vel_z": [-1.0, 1.0]}
    fused._set_program()
    assert P.command[0].lo[0] == 0.0 and P.command[0].hi[0] == 2.0
    env.reward_manager.cfg["action_rate"].weight = 0
    fused._set_program()
    assert P.reward[4].weight == 0.0
    env.velocity_command.use_external_controller(lambda step: env.velocity_command._command)
    fused._set_program()
    assert P.command[0].enabled == 0


def test_tilt_threshold_is_the_exact_decision_boundary():
    for deg in (10.0, 20.0, 30.0, 40.0):
        t = asin_tilt_threshold(deg)
        limit = math.radians(deg)
        below = np.nextafter(np.float32(t), np.float32(0.0))
        fires = lambda x: bool((torch.asin(torch.clamp(torch.tensor([x], dtype=torch.float32), max=0.99)) > limit).item())
        assert fires(t) and not fires(float(below))
        assert abs(t - math.sin(limit)) < 1e-6
    assert asin_tilt_threshold(85.0) == float("inf")  # asin(0.99) = 81.9 deg never exceeds 85


def test_user_defined_terms_select_split_execution(cpu_device):
    """Python reward / termination / observation terms and overridden command managers are kept as
    host callbacks (opcode EXTERNAL), and the step runs its kernel phases around them."""
    env = dry_env("custom_terms")
    fused, P, K = env._fused, env._fused.program.head, nat.K
    assert fused.split_mode
    names = [n for n, _, _ in fused.reward_terms]
    assert P.reward[names.index("speed")].op == K["GFB_R_EXTERNAL"]
    assert P.reward[names.index("contact_z")].ext_col == fused.ext_row[("reward", "contact_z")]
    tnames = [n for n, _, _ in fused.termination_terms]
    assert P.termination[tnames.index("wandered_off")].op == K["GFB_T_EXTERNAL"]
    assert [type(m).__name__ for m in fused.python_commands] == ["PythonCommand"]
    wave = fused.commands.index(fused.python_commands[0])
    assert P.command[wave].enabled == 0 and P.command[0].enabled == 1
    # external observation columns: 'clock' (2 wide) and the uncached angular velocity (3 wide)
    assert fused.ext_obs_width == 5 and fused.ext_obs.shape == (32, 5)
    srcs = [fused.program.obs_cols[c].src for c in range(P.obs_group[0].n_cols)]
    assert srcs.count(K["GFB_O_EXTERNAL"]) == 5
    assert not dry_env("contacts")._fused.split_mode


def test_no_cpu_fallback(cpu_device):
    spec = specs.get("simple")
    env = build_env(spec, dropin_namespace(), 8, torch.device("cpu"))
    with pytest.raises(nat.NativeLibraryError, match="no CPU implementation"):
        env.build()


def test_command_manager_surface(cpu_device):
    env = dry_env("kitchen_sink")
    vc: VelocityCommandManager = env.velocity_command
    assert vc.command.shape == (32, 3)
    assert vc.get_command_idx("ang_vel_z") == 2 and vc.get_command("lin_vel_x").shape == (32,)
    with pytest.raises(ValueError):
        vc.range = {"lin_vel_x": [0, 1]}
    with pytest.raises(ValueError):
        vc.range = (0.0, 1.0)
    hc: CommandManager = env.height_command
    with pytest.raises(ValueError):
        hc.get_command("x")
    vc.resample_time_sec = 2.0
    assert vc._resample_steps == 100
    assert vc.standing_probability == 0.02  # stored, no effect (dead code in the reference)


def test_contact_manager_surface(cpu_device):
    env = dry_env("contacts")
    cm = env.foot_contact_manager
    assert cm.link_ids.tolist() == [4, 8, 12, 16] and cm.local_link_ids.tolist() == [3, 7, 11, 15]
    assert cm.contacts.shape == (32, 4, 3) and cm.last_air_time.shape == (32, 4)
    cm._air[3, 0, 0] = 0.02
    assert bool(cm.has_made_contact(0.02)[0, 0]) and not bool(cm.has_made_contact(0.02)[0, 1])
    env2 = dry_env("berkeley_humanoid")
    with pytest.raises(RuntimeError, match="air time"):
        env2.torso_contact_manager.has_made_contact(0.02)


def test_reward_manager_surface(cpu_device):
    env = dry_env("command_direction")
    rm: RewardManager = env.reward_manager
    assert list(rm.episode_data) == list(specs.get("command_direction")["rewards"])
    assert all(v.shape == (32,) for v in rm.episode_data.values())
    assert rm.episode_data["lin_vel_z"].data_ptr() == rm._episode_sums[3].data_ptr()
    assert rm.last_episode_mean_reward("lin_vel_z") == 0.0


def test_term_descriptors_validate_arguments():
    with pytest.raises(AssertionError):
        rewards.command_tracking_lin_vel.gfb_signature(None)
    assert rewards.feet_air_time.gfb_opcode == "GFB_R_FEET_AIR_TIME"
    with pytest.raises(RuntimeError, match="not built"):
        rewards.is_alive(object())


def test_replay_rng():
    rng = ReplayRng()
    rng.push("a", torch.tensor([0.25, 0.5]))
    out = rng.uniform("a", torch.empty(2), -1, 1)
    assert out.tolist() == [0.25, 0.5]
    with pytest.raises(RuntimeError, match="no recorded draw"):
        rng.uniform("a", torch.empty(2), -1, 1)
    rng.push("b", torch.zeros(3))
    with pytest.raises(RuntimeError, match="shape"):
        rng.uniform("b", torch.empty(2), 0, 1)


def test_combine_logging_math():
    acc = torch.tensor([6.0, -3.0, 5.0, 0.0, 3.0], dtype=torch.float64)  # 2 rewards, 2 terminations, 3 resets
    out = combine_logging(acc, 2, 2, 100)
    assert out.tolist() == [2.0, -1.0, pytest.approx(0.05), 0.0]


def test_program_struct_is_plain_data():
    P = nat.Program()
    assert ctypes.sizeof(P) < 32 * 1024  # travels as a kernel parameter block / one memcpy
    assert ctypes.sizeof(nat.ProgramHead) % 8 == 0


# -- slab plans (host-only handle: gfb_spec_describe) -------------------------------------------------
def _plan(env, phases=None):
    from genesis_forge_b200 import spec

    fused = env._fused
    env._allocate_action_buffers(fused.D)
    fused.bind_action_buffers()
    fused.prepare_describe(injected=False)
    canon, words, tile = spec.describe(fused, nat.K["GFB_PHASE_ALL"] if phases is None else phases)
    names = ["tile", "needs", "n_staged", "n_early", "sums_late"]
    return dict(zip(names, words[:5]), smem_words=words[-1]), canon, words


def test_contact_tables_load_the_slab_in_two_groups(cpu_device, monkeypatch):
    """Arrays read only after the contact phase are loaded later, over the contact slots."""
    plain, _, _ = _plan(dry_env("command_direction", n=65536))
    assert plain["n_early"] == plain["n_staged"] and plain["sums_late"] == 0
    two, _, _ = _plan(dry_env("berkeley_humanoid", n=65536))
    assert 0 < two["n_early"] < two["n_staged"] and two["sums_late"] == 1
    monkeypatch.setenv("GFB_NO_OVERLAY", "1")
    one, _, _ = _plan(dry_env("berkeley_humanoid", n=65536))
    assert one["n_early"] == one["n_staged"] and one["sums_late"] == 0
    assert two["smem_words"] / two["tile"] < 0.8 * one["smem_words"] / one["tile"]  # shared memory per env


def test_plan_is_a_function_of_structure_not_of_live_values(cpu_device):
    """The specialisation key ignores weights / thresholds / step index, but not a zeroed weight's neighbours' layout."""
    env = dry_env("command_direction", n=65536)
    _, canon_a, words_a = _plan(env)
    env.reward_manager.cfg["lin_vel_z"].weight = -3.0
    env.step_count = 17
    env._fused._set_program(force=True)
    _, canon_b, words_b = _plan(env)
    assert canon_a == canon_b and words_a == words_b
    env.observation_managers["policy"].cfg["dof_position"].scale = 0.5   # a live value too
    env._fused._set_program(force=True)
    _, canon_c, words_c = _plan(env)
    assert canon_a == canon_c and words_a == words_c


def test_spawn_request_block(cpu_device):
    """gfb_spawn as TerrainManager fills it (terrain_manager.py:204-229: usable centre of the (sub)terrain)."""
    env = dry_env("rough_terrain")
    terrain = env.managers["terrain"][0]
    cfg = terrain._spawn_config(0.5, None, 0.4, {"z": (0.0, 2 * math.pi)})
    x_min, x_max, y_min, y_max = terrain.get_bounds()
    assert cfg.x_lo == np.float32(x_min + (x_max - x_min) / 4) and cfg.x_span == np.float32((x_max - x_min) / 2)
    assert cfg.y_lo == np.float32(y_min + (y_max - y_min) / 4) and cfg.y_span == np.float32((y_max - y_min) / 2)
    assert cfg.height_offset == np.float32(0.4) and cfg.with_rotation == 1
    assert list(cfg.rot_mode) == [0, 0, nat.K["GFB_SPAWN_ROT_DRAW"]]
    assert cfg.rot_hi[2] == np.float32(2 * math.pi) and cfg.rot_lo[2] == 0.0
    assert (cfg.height_field_rows, cfg.height_field_cols) == tuple(terrain.height_field.shape)
    assert terrain._spawn_config(0.5, None, 0.4, {"z": (0.0, 2 * math.pi)}) is cfg          # cached
    assert terrain._spawn_config(0.5, None, 0.4, None).with_rotation == 0                    # positions only
    with pytest.raises(nat.NativeLibraryError):
        terrain.generate_random_env_pos()  # no CUDA device here: loud failure, no host fallback


def test_product_code_never_imports_the_oracle():
    """oracle/ is test infrastructure: nothing under genesis_forge_b200/ may import or execute it."""
    import pathlib
    import re

    root = pathlib.Path(gfb.__file__).resolve().parent
    offenders = []
    for path in root.rglob("*.py"):
        text = path.read_text()
        if re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M) or "/root/reference" in text:
            offenders.append(str(path.relative_to(root)))
    assert offenders == []
