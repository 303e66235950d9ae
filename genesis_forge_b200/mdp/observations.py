"""
Observation terms, by the reference's names and signatures (genesis_forge/mdp/observations.py).
They are thin getters over the managers; inside an ObservationManager config the manager getters
hand out traced placeholders (see managers/observation.py), so these resolve to kernel columns.
"""
from __future__ import annotations

from .. import utils


def entity_linear_velocity(env, entity_manager=None, entity_attr="robot"):
    if entity_manager is not None:
        return entity_manager.get_linear_velocity()
    return env._trace_or("lin_vel_uncached", lambda: utils.entity_lin_vel(getattr(env, entity_attr)))


def entity_angular_velocity(env, entity_manager=None, entity_attr="robot"):
    if entity_manager is not None:
        return entity_manager.get_angular_velocity()
    return env._trace_or("ang_vel_uncached", lambda: utils.entity_ang_vel(getattr(env, entity_attr)))


def entity_projected_gravity(env, entity_manager=None, entity_attr="robot"):
    if entity_manager is not None:
        return entity_manager.get_projected_gravity()
    return env._trace_or("gravity_uncached", lambda: utils.entity_projected_gravity(getattr(env, entity_attr)))


def entity_dofs_position(env, action_manager=None, entity_attr="robot", dofs_idx=None):
    if action_manager is not None:
        return action_manager.get_dofs_position()
    return getattr(env, entity_attr).get_dofs_position(dofs_idx)


def entity_dofs_velocity(env, action_manager=None, entity_attr="robot", dofs_idx=None):
    if action_manager is not None:
        return action_manager.get_dofs_velocity()
    return getattr(env, entity_attr).get_dofs_velocity(dofs_idx)


def entity_dofs_force(env, action_manager=None, entity_attr="robot", dofs_idx=None, clip_to_max_force=False):
    if action_manager is not None:
        return action_manager.get_dofs_force(clip_to_max_force=clip_to_max_force)
    return getattr(env, entity_attr).get_dofs_force(dofs_idx)


def current_actions(env, action_manager=None):
    """The processed actions of this step (with an action manager) or the raw env actions."""
    if action_manager is not None:
        return action_manager.get_actions()
    return env._trace_or("env_actions", lambda: env.actions)


def contact_force(env, contact_manager):
    """Per tracked link |net contact force| (observations.py:181-193)."""
    return env._trace_or(("contact_norm", contact_manager), lambda: contact_manager.contacts.norm(dim=-1))
