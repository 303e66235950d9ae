"""
Reward terms, by the reference's names and signatures (genesis_forge/mdp/rewards.py).  Each function
body only normalises its arguments into the parameter dict the term compiler reads; the arithmetic
lives in csrc/post_kernel.cuh under the opcode named in the decorator.
"""
from __future__ import annotations

from ..managers.config import MdpFnClass
from ._term import term


@term("reward", "GFB_R_IS_ALIVE")
def is_alive(env):
    """1 for envs that did not terminate this step (rewards.py:31-37)."""
    return {}


@term("reward", "GFB_R_TERMINATED")
def terminated(env):
    """1 for envs that terminated (not timed out) this step (rewards.py:40-46)."""
    return {}


@term("reward", "GFB_R_BASE_HEIGHT")
def base_height(env, target_height=None, height_command=None, terrain_manager=None, entity_attr="robot",
                entity_manager=None):
    """(base z - terrain height - target)^2 (rewards.py:54-90)."""
    return dict(target_height=target_height, height_command=height_command, terrain_manager=terrain_manager,
                entity_attr=entity_attr, entity_manager=entity_manager)


@term("reward", "GFB_R_DOF_SIMILAR")
def dof_similar_to_default(env, action_manager):
    """sum_d |q_d - default_d| (rewards.py:93-109)."""
    return dict(action_manager=action_manager)


@term("reward", "GFB_R_LIN_VEL_Z")
def lin_vel_z_l2(env, entity_attr="robot", entity_manager=None):
    """Squared body-frame vertical velocity (rewards.py:112-135)."""
    return dict(entity_attr=entity_attr, entity_manager=entity_manager)


@term("reward", "GFB_R_ANG_VEL_XY")
def ang_vel_xy_l2(env, entity_attr="robot", entity_manager=None):
    """Squared body-frame roll/pitch rates (rewards.py:138-161)."""
    return dict(entity_attr=entity_attr, entity_manager=entity_manager)


@term("reward", "GFB_R_FLAT_ORIENTATION")
def flat_orientation_l2(env, entity_attr="robot", entity_manager=None):
    """Squared xy of the projected gravity (rewards.py:164-193)."""
    return dict(entity_attr=entity_attr, entity_manager=entity_manager)


@term("reward", "GFB_R_ACTION_RATE")
def action_rate_l2(env):
    """sum_d (last_action_d - action_d)^2 on the raw env actions (rewards.py:257-271)."""
    return {}


@term("reward", "GFB_R_TRACK_LIN_VEL")
def command_tracking_lin_vel(env, command=None, vel_cmd_manager=None, sensitivity=0.25, entity_attr="robot",
                             entity_manager=None):
    """exp(-|cmd_xy - v_xy|^2 / sensitivity) (rewards.py:279-317)."""
    assert command is not None or vel_cmd_manager is not None, (
        "Either command or vel_cmd_manager must be provided to command_tracking_lin_vel"
    )
    return dict(command=command, vel_cmd_manager=vel_cmd_manager, sensitivity=sensitivity,
                entity_attr=entity_attr, entity_manager=entity_manager)


@term("reward", "GFB_R_TRACK_ANG_VEL")
def command_tracking_ang_vel(env, commanded_ang_vel=None, vel_cmd_manager=None, sensitivity=0.25,
                             entity_attr="robot", entity_manager=None):
    """exp(-(cmd_yaw - w_z)^2 / sensitivity) (rewards.py:320-358)."""
    assert commanded_ang_vel is not None or vel_cmd_manager is not None, (
        "Either commanded_ang_vel or vel_cmd_manager must be provided to command_tracking_ang_vel"
    )
    return dict(commanded_ang_vel=commanded_ang_vel, vel_cmd_manager=vel_cmd_manager, sensitivity=sensitivity,
                entity_attr=entity_attr, entity_manager=entity_manager)


@term("reward", "GFB_R_STAND_STILL")
def stand_still_joint_deviation_l1(env, command_threshold=0.06, vel_cmd_manager=None, action_manager=None):
    """Joint deviation from default while the xy command is below a threshold (rewards.py:361-385)."""
    return dict(command_threshold=command_threshold, vel_cmd_manager=vel_cmd_manager, action_manager=action_manager)


@term("reward", "GFB_R_HAS_CONTACT")
def has_contact(_env, contact_manager, threshold=1.0, min_contacts=1):
    """1 when at least `min_contacts` tracked links exceed the force threshold (rewards.py:393-410)."""
    return dict(contact_manager=contact_manager, threshold=threshold, min_contacts=min_contacts)


@term("reward", "GFB_R_CONTACT_FORCE")
def contact_force(_env, contact_manager, threshold=1.0):
    """Total contact force above the threshold over the tracked links (rewards.py:413-428)."""
    return dict(contact_manager=contact_manager, threshold=threshold)


@term("reward", "GFB_R_FEET_AIR_TIME")
def feet_air_time(env, contact_manager, time_threshold, time_threshold_max=None, vel_cmd_manager=None):
    """Air time beyond a threshold, credited at touchdown, gated by the command (rewards.py:431-469)."""
    return dict(contact_manager=contact_manager, time_threshold=time_threshold,
                time_threshold_max=time_threshold_max, vel_cmd_manager=vel_cmd_manager)


@term("reward", "GFB_R_FEET_SLIDE")
def feet_slide(env, contact_manager, entity_attr="robot"):
    """Speed of tracked links while they are in contact (rewards.py:472-504)."""
    return dict(contact_manager=contact_manager, entity_attr=entity_attr)


class body_acceleration_exp(MdpFnClass):
    """
    Penalise jerky body motion: 1 - exp(-sensitivity * (|d v_body / dt| + |d w_body / dt|))
    (rewards.py:196-249).  A class-style term: it keeps the previous step's body-frame velocities,
    which are not cleared on reset; its first evaluation sees zero acceleration.  Evaluated by the
    kernel (opcode GFB_R_BODY_ACC_EXP); the state lives in a (N,6) tensor owned by the fused step.
    As in the reference an `entity_manager` is required (the reference's `entity_attr` path reads an
    attribute it never sets).
    """

    gfb_kind = "reward"
    gfb_opcode = "GFB_R_BODY_ACC_EXP"

    def __init__(self, env, entity_attr="robot", entity_manager=None, sensitivity=0.10):
        super().__init__(env)
        self.evaluations = 0

    @staticmethod
    def gfb_signature(env, entity_attr="robot", entity_manager=None, sensitivity=0.10):
        if entity_manager is None:
            raise AttributeError("'body_acceleration_exp' object has no attribute '_entity_attr'")
        return dict(entity_attr=entity_attr, entity_manager=entity_manager, sensitivity=sensitivity)

    def __call__(self, env, entity_attr="robot", entity_manager=None, sensitivity=0.10):
        return env._fused.evaluate_single_term("reward", self, self.gfb_signature(env, entity_attr, entity_manager, sensitivity))
