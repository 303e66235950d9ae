"""
Workload with TWO EntityManagers: the robot of a spec plus a free rigid body ("prop") with its own
manager, on_reset item and observation group.  The reference's registry keeps entity managers in a
list (genesis_forge/managed_env.py:200-220, :269-270, :294-296, :353-354); the drop-in's kernel serves
the first one and keeps the others on the host side of the step (managers/entity.py).

Like configs/env_builder.py this is namespace-agnostic: `add_prop(env, ns)` extends the `config()` of an
environment built from the reference's or from the drop-in's classes with the same manager table.
"""
from __future__ import annotations

import torch

PROP_RESET_POS = [0.5, -0.25, 0.1]
PROP_RESET_QUAT = [1.0, 0.0, 0.0, 0.0]


def spec(stock_terms: bool = False):
    """`stock_terms`: also stock mdp reward / termination terms that refer to the prop -- through its manager
    (cached pose) and through `entity_attr` (current pose)."""
    from . import specs

    s = specs.get("simple")
    s["name"] = "second_entity_terms" if stock_terms else "second_entity"
    s["max_episode_random_scaling"] = 0.0  # nothing in this workload depends on a random draw
    if stock_terms:
        prop = {"entity_manager": "@prop_manager"}
        s["rewards"].update({
            "prop_lin_vel_z": {"fn": "lin_vel_z_l2", "weight": -0.5, "params": dict(prop)},
            "prop_flat": {"fn": "flat_orientation_l2", "weight": -0.25, "params": dict(prop)},
            "prop_height": {"fn": "base_height", "weight": -2.0, "params": {"target_height": 1.3, "entity_attr": "prop"}},
        })
        s["terminations"]["prop_tilt"] = {"fn": "bad_orientation", "params": {"limit_angle": 4.5, **prop}}
    return s


class Prop:
    """
    Second entity of the synthetic engine.  Its base state is a fixed, exactly representable remix of the
    robot's synthetic state (the quaternion a permutation of the robot's, hence a unit quaternion too),
    regenerated once per physics step; setters apply to the current step's state like SyntheticRobot's.
    """

    def __init__(self, scene):
        self._scene = scene
        self._at = self._src = None
        self.calls: list[tuple] = []

    def _state(self) -> dict:
        scene = self._scene
        if self._at != scene.step_index or self._src is not scene.state:
            st = scene.state
            self._pos = st["pos"] + 1.0
            self._quat = st["quat"][:, [0, 2, 3, 1]].contiguous()
            self._vel = st["ang"] * 0.5
            self._ang = st["vel"] * 2.0
            self._at, self._src = scene.step_index, st
        return self

    def _out(self, t):
        return t.clone() if self._scene.copy_on_get else t

    def get_pos(self, envs_idx=None):
        return self._out(self._state()._pos)

    def get_quat(self, envs_idx=None):
        return self._out(self._state()._quat)

    def get_vel(self, envs_idx=None):
        return self._out(self._state()._vel)

    def get_ang(self, envs_idx=None):
        return self._out(self._state()._ang)

    def _zero_velocity(self, envs_idx):
        self._vel[envs_idx] = 0.0
        self._ang[envs_idx] = 0.0

    def set_pos(self, pos, envs_idx=None, zero_velocity=True, relative=False):
        self._state()._pos[envs_idx] = pos
        if zero_velocity:
            self._zero_velocity(envs_idx)
        self.calls.append(("set_pos", len(envs_idx)))

    def set_quat(self, quat, envs_idx=None, zero_velocity=True, relative=False):
        self._state()._quat[envs_idx] = quat
        if zero_velocity:
            self._zero_velocity(envs_idx)
        self.calls.append(("set_quat", len(envs_idx)))


def add_prop(env, ns):
    """Extend env.config(): a prop entity, its EntityManager (registered right after the robot's, before the
    term managers that may refer to it) and an observation group around its getters."""
    base_config = env.config

    def after_entity_manager():
        env.prop = Prop(env.scene)
        env.prop_manager = ns.managers.EntityManager(
            env, entity_attr="prop",
            on_reset={"position": {"fn": ns.reset.position,
                                   "params": {"position": PROP_RESET_POS, "quat": PROP_RESET_QUAT, "zero_velocity": True}}},
        )

    def config():
        base_config()
        env.observation_managers["prop"] = ns.managers.ObservationManager(
            env, name="prop",
            cfg={
                "prop_linear_velocity": {"fn": lambda env: env.prop_manager.get_linear_velocity(), "scale": 2.0},
                "prop_projected_gravity": {"fn": lambda env: env.prop_manager.get_projected_gravity()},
                "prop_angular_velocity": {"fn": lambda env: env.prop_manager.get_angular_velocity(), "scale": 0.25},
                # the robot's own manager inside the same group: served by the kernel
                "robot_angular_velocity": {"fn": lambda env: env.robot_manager.get_angular_velocity()},
            },
        )

    env.after_entity_manager = after_entity_manager
    env.config = config
    return env


def prop_cache(env) -> dict:
    """The prop manager's cached pose (what a user reads between steps)."""
    m = env.prop_manager
    return {"base_pos": m.base_pos.detach().cpu().clone(), "base_quat": m.base_quat.detach().cpu().clone(),
            "inv_base_quat": m.inv_base_quat.detach().cpu().clone()}
