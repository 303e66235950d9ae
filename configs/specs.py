"""
The BASELINE.json configs as term-table specs (workload definitions; no arithmetic of the path).

Each spec is a plain dict transcribed from the `config()` method of the reference example it names
(the examples ARE the benchmark definitions).  The same spec drives
  * configs/env_builder.py  -> a ManagedEnvironment built from the reference's (or the drop-in's)
                              own manager classes and mdp functions,
  * oracle/manager_port.py -> the torch-CPU oracle port.
Manager references inside params are written "@<manager name>".
"""
from __future__ import annotations

import copy

GO2_JOINTS = ["FL_.*_joint", "FR_.*_joint", "RL_.*_joint", "RR_.*_joint"]
GO2_DEFAULT_POS = {
    ".*_hip_joint": 0.0,
    "FL_thigh_joint": 0.8,
    "FR_thigh_joint": 0.8,
    "RL_thigh_joint": 1.0,
    "RR_thigh_joint": 1.0,
    ".*_calf_joint": -1.5,
}
INITIAL_POS = [0.0, 0.0, 0.4]
INITIAL_QUAT = [1.0, 0.0, 0.0, 0.0]

_RESET_FIXED = {
    "position": {
        "fn": "position",
        "params": {"position": INITIAL_POS, "quat": INITIAL_QUAT, "zero_velocity": True},
    }
}

_OBS_48 = {
    "velocity_cmd": {"fn": "command", "mgr": "velocity_command"},
    "angle_velocity": {"fn": "ang_vel"},
    "linear_velocity": {"fn": "lin_vel"},
    "projected_gravity": {"fn": "gravity"},
    "dof_position": {"fn": "dof_pos"},
    "dof_velocity": {"fn": "dof_vel", "scale": 0.05},
    "actions": {"fn": "actions"},
}

_EM = "@robot_manager"
_VC = "@velocity_command"
_AM = "@action_manager"


def _track(weight_lin, weight_ang):
    return {
        "tracking_lin_vel": {
            "fn": "command_tracking_lin_vel", "weight": weight_lin,
            "params": {"vel_cmd_manager": _VC, "entity_manager": _EM},
        },
        "tracking_ang_vel": {
            "fn": "command_tracking_ang_vel", "weight": weight_ang,
            "params": {"vel_cmd_manager": _VC, "entity_manager": _EM},
        },
    }


def simple() -> dict:
    """examples/simple/environment.py:96-243 (config 1; fixed target command, no command manager)."""
    return {
        "name": "simple", "robot": "go2", "dt": 1 / 50,
        "max_episode_length_sec": 20, "max_episode_random_scaling": 0.1,
        "fixed_command": [0.5, 0.0, 0.0],
        "entity": {"on_reset": _RESET_FIXED},
        "action": {
            "type": "position", "joint_names": GO2_JOINTS, "default_pos": GO2_DEFAULT_POS,
            "scale": 0.25, "clip": (-100.0, 100.0), "use_default_offset": True, "pd_kp": 20, "pd_kv": 0.5,
        },
        "commands": {},
        "contacts": {},
        "rewards": {
            "base_height_target": {
                "fn": "base_height", "weight": -50.0, "params": {"target_height": 0.3, "entity_attr": "robot"},
            },
            "tracking_lin_vel": {
                "fn": "command_tracking_lin_vel", "weight": 1.0,
                "params": {"command": "@fixed_command[:, :2]", "entity_manager": _EM},
            },
            "tracking_ang_vel": {
                "fn": "command_tracking_ang_vel", "weight": 0.2,
                "params": {"commanded_ang_vel": "@fixed_command[:, 2]", "entity_manager": _EM},
            },
            "lin_vel_z": {"fn": "lin_vel_z_l2", "weight": -1.0, "params": {"entity_manager": _EM}},
            "action_rate": {"fn": "action_rate_l2", "weight": -0.005},
            "similar_to_default": {"fn": "dof_similar_to_default", "weight": -0.1, "params": {"action_manager": _AM}},
        },
        "terminations": {
            "timeout": {"fn": "timeout", "time_out": True},
            "fall_over": {"fn": "bad_orientation", "params": {"limit_angle": 10.0, "entity_manager": _EM}},
        },
        "observations": {
            "policy": {
                "terms": {
                    "angle_velocity": {"fn": "ang_vel", "scale": 0.25},
                    "linear_velocity": {"fn": "lin_vel", "scale": 2.0},
                    "projected_gravity": {"fn": "gravity"},
                    "dof_position": {"fn": "dof_pos"},
                    "dof_velocity": {"fn": "dof_vel", "scale": 0.05},
                    "actions": {"fn": "actions"},
                }
            }
        },
    }


def command_direction() -> dict:
    """examples/command_direction/environment.py:85-247 (config 2)."""
    return {
        "name": "command_direction", "robot": "go2", "dt": 1 / 50,
        "max_episode_length_sec": 20, "max_episode_random_scaling": 0.1,
        "entity": {"on_reset": _RESET_FIXED},
        "action": {
            "type": "position", "joint_names": GO2_JOINTS, "default_pos": GO2_DEFAULT_POS,
            "scale": 0.25, "use_default_offset": True, "pd_kp": 20, "pd_kv": 0.5,
        },
        "commands": {
            "velocity_command": {
                "type": "velocity",
                "range": {"lin_vel_x": [-1.0, 1.0], "lin_vel_y": [-1.0, 1.0], "ang_vel_z": [-1.0, 1.0]},
                "standing_probability": 0.02, "resample_time_sec": 5.0,
            }
        },
        "contacts": {},
        "rewards": {
            "base_height_target": {
                "fn": "base_height", "weight": -50.0, "params": {"target_height": 0.3, "entity_attr": "robot"},
            },
            **_track(1.0, 0.5),
            "lin_vel_z": {"fn": "lin_vel_z_l2", "weight": -1.0, "params": {"entity_manager": _EM}},
            "action_rate": {"fn": "action_rate_l2", "weight": -0.005},
            "similar_to_default": {"fn": "dof_similar_to_default", "weight": -0.1, "params": {"action_manager": _AM}},
        },
        "terminations": {
            "timeout": {"fn": "timeout", "time_out": True},
            "fall_over": {"fn": "bad_orientation", "params": {"limit_angle": 10.0, "entity_manager": _EM}},
        },
        "observations": {"policy": {"terms": copy.deepcopy(_OBS_48)}},
    }


def contacts() -> dict:
    """examples/contacts/environment.py:89-272 (config 3a)."""
    return {
        "name": "contacts", "robot": "go2", "dt": 1 / 50,
        "max_episode_length_sec": 20, "max_episode_random_scaling": 0.1,
        "entity": {"on_reset": _RESET_FIXED},
        "action": {
            "type": "position", "joint_names": GO2_JOINTS, "default_pos": GO2_DEFAULT_POS,
            "scale": 0.5, "use_default_offset": True, "pd_kp": 20, "pd_kv": 0.5,
        },
        "commands": {
            "velocity_command": {
                "type": "velocity",
                "range": {"lin_vel_x": [-1.0, 1.0], "lin_vel_y": [0.0, 0.0], "ang_vel_z": [-0.5, 0.5]},
                "standing_probability": 0.02, "resample_time_sec": 5.0,
            }
        },
        "contacts": {
            "foot_contact_manager": {
                "link_names": [".*_calf"], "track_air_time": True, "air_time_contact_threshold": 5.0,
            }
        },
        "rewards": {
            "foot_air_time": {
                "fn": "feet_air_time", "weight": 2.5,
                "params": {"contact_manager": "@foot_contact_manager", "vel_cmd_manager": _VC, "time_threshold": 0.5},
            },
            **_track(1.0, 0.5),
            "lin_vel_z": {"fn": "lin_vel_z_l2", "weight": -1.0, "params": {"entity_manager": _EM}},
            "ang_vel_xy": {"fn": "ang_vel_xy_l2", "weight": -0.05, "params": {"entity_manager": _EM}},
            "action_rate": {"fn": "action_rate_l2", "weight": -0.005},
            "similar_to_default": {"fn": "dof_similar_to_default", "weight": -0.1, "params": {"action_manager": _AM}},
            "flat_orientation": {"fn": "flat_orientation_l2", "weight": -2.5},
        },
        "terminations": {
            "timeout": {"fn": "timeout", "time_out": True},
            "fall_over": {"fn": "bad_orientation", "params": {"limit_angle": 20.0, "entity_manager": _EM}},
        },
        "observations": {"policy": {"terms": copy.deepcopy(_OBS_48)}},
    }


def gait_trainer() -> dict:
    """
    examples/gait_trainer/environment.py:94-336 (config 3b): three ContactManagers, the velocity
    command plus the example's own GaitCommandManager (gait clock, two user-defined reward terms,
    examples/gait_trainer/gait_command_manager.py:224-345), policy (H=5, O=62) and critic (H=5, O=16)
    observation groups.  `curriculum` advances the manager's curriculum state before the run the way
    Go2GaitTrainingEnv.update_curriculum (:359-389) does during training, so that several gaits and
    non-degenerate period / clearance ranges are sampled.
    """
    return {
        "name": "gait_trainer", "robot": "go2", "dt": 1 / 50,
        "max_episode_length_sec": 20, "max_episode_random_scaling": 0.4,
        "entity": {"on_reset": _RESET_FIXED},
        "action": {
            "type": "position", "joint_names": GO2_JOINTS, "default_pos": GO2_DEFAULT_POS,
            "scale": 0.25, "use_default_offset": True, "pd_kp": 20, "pd_kv": 0.5,
        },
        "commands": {
            "velocity_command": {
                "type": "velocity",
                "range": {"lin_vel_x": [-1.0, 1.0], "lin_vel_y": [0.0, 0.0], "ang_vel_z": [-1.0, 1.0]},
                "standing_probability": 0.0, "resample_time_sec": 3.0,
            },
            "gait_command_manager": {
                "type": "gait", "resample_time_sec": 4.0,
                "foot_names": {"FL": "FL_foot", "FR": "FR_foot", "RL": "RL_foot", "RR": "RR_foot"},
                "curriculum": {"num_gaits": 2, "gait_period_range": 2, "foot_clearance_range": 3},
            },
        },
        "contacts": {
            "foot_contact_manager": {"link_names": [".*_foot"], "air_time_contact_threshold": 1.0},
            "body_contact_manager": {"link_names": ["base"], "air_time_contact_threshold": 1.0},
            "bad_contact_manager": {"link_names": [".*_thigh", ".*_calf"]},
        },
        "rewards": {
            "gait_phase_reward": {
                "fn": "@gait_command_manager.gait_phase_reward", "weight": 1.5,
                "params": {"contact_manager": "@foot_contact_manager"},
            },
            "foot_height_reward": {"fn": "@gait_command_manager.foot_height_reward", "weight": 0.9},
            "base_height_target": {
                "fn": "base_height", "weight": -25.0, "params": {"target_height": 0.35, "entity_attr": "robot"},
            },
            **_track(1.0, 0.5),
            "body_acceleration": {"fn": "body_acceleration_exp", "weight": -0.1, "params": {"entity_manager": _EM}},
            "lin_vel_z": {"fn": "lin_vel_z_l2", "weight": -0.1, "params": {"entity_manager": _EM}},
            "action_rate": {"fn": "action_rate_l2", "weight": -0.01},
            "bad_contact": {"fn": "contact_force", "weight": -1.0, "params": {"contact_manager": "@bad_contact_manager"}},
        },
        "terminations": {
            "timeout": {"fn": "timeout", "time_out": True},
            "fall_over": {"fn": "bad_orientation", "params": {"limit_angle": 20.0, "entity_manager": _EM}},
            "body_contact": {
                "fn": "contact_force", "params": {"contact_manager": "@body_contact_manager", "threshold": 1.0},
            },
        },
        "observations": {
            "policy": {
                "history_len": 5,
                "terms": {"gait_command": {"fn": "command", "mgr": "gait_command_manager"}, **copy.deepcopy(_OBS_48)},
            },
            "critic": {
                "history_len": 5,
                "terms": {
                    "foot_contact_force": {"fn": "contact_force", "mgr": "foot_contact_manager"},
                    "dof_force": {"fn": "entity_dofs_force", "scale": 0.1},
                },
            },
        },
    }


def rough_terrain() -> dict:
    """examples/rough_terrain/environment.py:86-321 (config 4)."""
    return {
        "name": "rough_terrain", "robot": "go2", "dt": 1 / 50,
        "max_episode_length_sec": 20, "max_episode_random_scaling": 0.1,
        "terrain": {
            "pos": (-12.0, -12.0, 0.0), "n_subterrains": (1, 1), "subterrain_size": (24.0, 24.0),
            "vertical_scale": 0.001,
        },
        "xy_range": 11.6,  # terrain spans +-12 m, out_of_bounds margin 0.5 -> ~1.7 % of envs per step
        "entity": {
            "on_reset": {
                "position": {
                    "fn": "randomize_terrain_position",
                    "params": {"height_offset": 0.4, "terrain_manager": "@terrain_manager"},
                }
            }
        },
        "action": {
            "type": "position", "joint_names": GO2_JOINTS, "default_pos": GO2_DEFAULT_POS,
            "scale": 0.25, "use_default_offset": True, "pd_kp": 20, "pd_kv": 0.5, "max_force": 23.5,
        },
        "commands": {
            "velocity_command": {
                "type": "velocity",
                "range": {"lin_vel_x": [-1.0, 1.0], "lin_vel_y": [-1.0, 1.0], "ang_vel_z": [-0.5, 0.5]},
                "standing_probability": 0.05, "resample_time_sec": 5.0,
            }
        },
        "contacts": {
            "foot_contact_manager": {
                "link_names": [".*_calf"], "track_air_time": True, "air_time_contact_threshold": 5.0,
            },
            "undesired_contacts": {"link_names": [".*_thigh", "base"]},
        },
        "rewards": {
            **_track(1.5, 0.75),
            "lin_vel_z": {"fn": "lin_vel_z_l2", "weight": -2.0, "params": {"entity_manager": _EM}},
            "ang_vel_xy": {"fn": "ang_vel_xy_l2", "weight": -0.05, "params": {"entity_manager": _EM}},
            "undesired_contacts": {
                "fn": "has_contact", "weight": -1.0,
                "params": {"contact_manager": "@undesired_contacts", "threshold": 5.0},
            },
            "action_rate": {"fn": "action_rate_l2", "weight": -0.01},
            "similar_to_default": {"fn": "dof_similar_to_default", "weight": -0.1, "params": {"action_manager": _AM}},
            "flat_orientation": {"fn": "flat_orientation_l2", "weight": -1.5},
            "terminated": {"fn": "terminated", "weight": -100.0},
        },
        "terminations": {
            "timeout": {"fn": "timeout", "time_out": True},
            "out_of_bounds": {"fn": "out_of_bounds", "params": {"terrain_manager": "@terrain_manager"}},
            "bad_orientation": {
                "fn": "bad_orientation",
                "params": {"limit_angle": 30.0, "entity_manager": _EM, "grace_steps": 20},
            },
        },
        "observations": {"policy": {"terms": copy.deepcopy(_OBS_48)}},
    }


def berkeley_humanoid() -> dict:
    """examples/berkeley_humanoid/environment.py:85-274 (config 5)."""
    return {
        "name": "berkeley_humanoid", "robot": "berkeley_humanoid", "dt": 1 / 50,
        "max_episode_length_sec": 20, "max_episode_random_scaling": 0.1,
        "entity": {
            "on_reset": {
                "position": {
                    "fn": "position",
                    "params": {"position": [0.0, 0.0, 0.55], "quat": INITIAL_QUAT, "zero_velocity": True},
                }
            }
        },
        "action": {
            "type": "position", "joint_names": [".*"],
            "default_pos": {
                "LL_HR": -0.071, "LR_HR": 0.071, "LL_HAA": 0.103, "LR_HAA": -0.103,
                "LL_HFE": -0.463, "LR_HFE": -0.463, "LL_KFE": 0.983, "LR_KFE": 0.983,
                "LL_FFE": -0.350, "LR_FFE": -0.350, "LL_FAA": 0.126, "LR_FAA": -0.126,
            },
            "scale": 0.5, "use_default_offset": True, "pd_kp": 15.0, "pd_kv": 1.0,
            "max_force": {".*_HR": 20.0, ".*_HAA": 20.0, ".*_HFE": 30.0, ".*_KFE": 30.0, ".*_FFE": 20.0, ".*_FAA": 5.0},
        },
        "commands": {
            "velocity_command": {
                "type": "velocity",
                "range": {"lin_vel_x": [0.0, 1.0], "lin_vel_y": [0.0, 0.0], "ang_vel_z": [-0.5, 0.5]},
                "standing_probability": 0.02, "resample_time_sec": 5.0,
            }
        },
        "contacts": {
            "torso_contact_manager": {"link_names": ["torso"]},
            "feet_contact_manager": {"link_names": [".*_faa"], "track_air_time": True},
        },
        "rewards": {
            **_track(1.0, 0.5),
            "lin_vel_z": {"fn": "lin_vel_z_l2", "weight": -2.0, "params": {"entity_manager": _EM}},
            "ang_vel_xy_l2": {"fn": "ang_vel_xy_l2", "weight": -0.05, "params": {"entity_manager": _EM}},
            "action_rate": {"fn": "action_rate_l2", "weight": -0.005},
            "similar_to_default": {"fn": "dof_similar_to_default", "weight": -0.05, "params": {"action_manager": _AM}},
            "feet_air_time": {
                "fn": "feet_air_time", "weight": 2.0,
                "params": {
                    "time_threshold": 0.2, "time_threshold_max": 0.5,
                    "contact_manager": "@feet_contact_manager", "vel_cmd_manager": _VC,
                },
            },
        },
        "terminations": {
            "timeout": {"fn": "timeout", "time_out": True},
            "torso_contact": {"fn": "contact_force", "params": {"contact_manager": "@torso_contact_manager"}},
        },
        "observations": {"policy": {"terms": copy.deepcopy(_OBS_48)}},
    }


def kitchen_sink() -> dict:
    """
    Not a BASELINE config: a Go2 table that exercises every remaining mdp term of SURVEY.md 8(a)
    (is_alive, stand_still, contact_force reward, feet_slide, base_height_below_minimum, has_contact
    / contact_force_with_grace_period terminations, observation noise + history, dof_force and
    contact-force observations, a with-entity contact filter, a height command, action delay).
    """
    s = contacts()
    s["name"] = "kitchen_sink"
    s["action"]["delay_step"] = 2
    s["action"]["noise_scale"] = 0.05
    s["commands"]["height_command"] = {"type": "uniform", "range": (0.25, 0.35), "resample_time_sec": 3.0}
    s["contacts"]["thigh_ground"] = {"link_names": [".*_thigh"], "with_entity_attr": "terrain"}
    s["contacts"]["body"] = {"link_names": ["base", ".*_hip"], "with_links_names": [".*_calf", ".*_foot"]}
    s["rewards"].update({
        "alive": {"fn": "is_alive", "weight": 0.5},
        "height_cmd": {
            "fn": "base_height", "weight": -10.0,
            "params": {"height_command": "@height_command", "entity_manager": _EM},
        },
        "stand_still": {
            "fn": "stand_still_joint_deviation_l1", "weight": -0.2,
            "params": {"command_threshold": 0.3, "vel_cmd_manager": _VC, "action_manager": _AM},
        },
        "thigh_force": {
            "fn": "contact_force", "weight": -0.01,
            "params": {"contact_manager": "@thigh_ground", "threshold": 2.0},
        },
        "feet_slide": {"fn": "feet_slide", "weight": -0.1, "params": {"contact_manager": "@foot_contact_manager"}},
        "zero_weight": {"fn": "lin_vel_z_l2", "weight": 0.0, "params": {"entity_manager": _EM}},
        "body_acceleration": {
            "fn": "body_acceleration_exp", "weight": -0.1,
            "params": {"entity_manager": _EM, "sensitivity": 0.15},
        },
    })
    s["terminations"].update({
        "too_low": {"fn": "base_height_below_minimum", "params": {"minimum_height": 0.24, "entity_manager": _EM}},
        "body_hit": {
            "fn": "has_contact", "params": {"contact_manager": "@body", "threshold": 40.0, "min_contacts": 2},
        },
        "thigh_hit": {
            "fn": "contact_force_with_grace_period",
            "params": {"contact_manager": "@thigh_ground", "threshold": 90.0, "grace_steps": 5},
        },
    })
    terms = s["observations"]["policy"]["terms"]
    terms["angle_velocity"]["noise"] = 0.1
    terms["dof_position"]["noise"] = 0.01
    terms["dof_force"] = {"fn": "dof_force", "scale": 0.1}
    terms["foot_force"] = {"fn": "contact_force", "mgr": "foot_contact_manager", "scale": 0.01}
    terms["height_cmd"] = {"fn": "command", "mgr": "height_command"}
    # (mdp.observations.current_actions without an action manager cannot be used: env.actions is
    #  still None during ObservationManager.build()'s dry run in the reference.)
    terms["actions2"] = {"fn": "current_actions"}
    s["observations"]["policy"]["history_len"] = 3
    s["observations"]["critic"] = {
        "noise": 0.02,
        "terms": {
            "linear_velocity": {"fn": "lin_vel"},
            "projected_gravity": {"fn": "gravity", "noise": 0.05},
            "dof_velocity": {"fn": "dof_vel", "scale": 0.05, "noise": 0.0},
        },
    }
    return s


def custom_terms() -> dict:
    """
    Not a BASELINE config: user-defined (Python) reward / termination / observation terms and a
    command manager with overridden step()/reset(), the way examples/gait_trainer extends the
    library (gait_command_manager.py).  The callables only use API that the reference and the
    drop-in share, so the same objects drive the reference, the port and the CUDA path (which runs
    them as host callbacks between kernel phases).
    """
    import torch

    s = contacts()
    s["name"] = "custom_terms"

    def speed_reward(env):
        return env.robot.get_vel()[:, 0].abs()

    def contact_reward(env, gain=0.01):
        return env.foot_contact_manager.contacts[:, :, 2].sum(dim=1) * gain + env.extras["terminations"].float()

    def wandered_off(env, limit=9.5):
        return env.robot.get_pos()[:, 0] > limit

    def clock_obs(env):
        return torch.stack([env.episode_length.float() * 0.001, env.robot.get_pos()[:, 2]], dim=1)

    def wave_step(mgr, env):
        mgr._command[:, 0] = 0.5 + 0.001 * float(env.step_count % 7)

    def wave_reset(mgr, env, env_ids):
        if env_ids is None:
            mgr._command[:, 0] = -1.0
        else:
            mgr._command[env_ids, 0] = -1.0

    s["commands"]["wave"] = {"type": "python", "range": (-1.0, 1.0), "resample_time_sec": 5.0,
                             "step": wave_step, "reset": wave_reset}
    s["rewards"]["speed"] = {"fn": speed_reward, "weight": 0.3}
    s["rewards"]["contact_z"] = {"fn": contact_reward, "weight": -0.2, "params": {"gain": 0.02}}
    s["rewards"]["body_acceleration"] = {
        "fn": "body_acceleration_exp", "weight": -0.1, "params": {"entity_manager": _EM},
    }
    s["terminations"]["wandered_off"] = {"fn": wandered_off, "params": {"limit": 9.6}}
    terms = s["observations"]["policy"]["terms"]
    terms["clock"] = {"fn": clock_obs, "scale": 2.0}
    terms["wave"] = {"fn": "command", "mgr": "wave"}
    terms["ang_vel_fresh"] = {"fn": "ang_vel_uncached", "noise": 0.05}
    return s


ALL = {
    "simple": simple,
    "command_direction": command_direction,
    "contacts": contacts,
    "gait_trainer": gait_trainer,
    "rough_terrain": rough_terrain,
    "berkeley_humanoid": berkeley_humanoid,
    "kitchen_sink": kitchen_sink,
    "custom_terms": custom_terms,
}


def within_limits() -> dict:
    """
    Variant of config 2 with the PositionWithinLimitsActionManager (SURVEY.md 8(a) row a4:
    actions clamped to [-1, 1] and mapped onto each joint's limits, no NaN/Inf check).
    """
    s = command_direction()
    s["name"] = "within_limits"
    action = dict(s["action"], type="within_limits")
    for key in ("scale", "use_default_offset"):
        action.pop(key, None)
    s["action"] = action
    return s


# variants of the configs above: golden-traced and parity-tested like them, but no specialised
# kernels are pre-built for them (tools/prebuild_specs.py walks ALL)
VARIANTS = {
    "within_limits": within_limits,
}


def get(name: str) -> dict:
    return (ALL.get(name) or VARIANTS[name])()
