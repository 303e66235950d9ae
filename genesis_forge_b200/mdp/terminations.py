"""
Termination terms, by the reference's names and signatures (genesis_forge/mdp/terminations.py).
Descriptors for the opcodes evaluated in csrc/post_kernel.cuh (see mdp/_term.py).
"""
from __future__ import annotations

from ._term import term


@term("termination", "GFB_T_TIMEOUT")
def timeout(env):
    """episode_length > max_episode_length (terminations.py:17-23)."""
    return {}


@term("termination", "GFB_T_BAD_ORIENTATION")
def bad_orientation(env, limit_angle=40.0, entity_attr="robot", entity_manager=None, grace_steps=0):
    """Tilt angle from the projected gravity exceeds `limit_angle` degrees (terminations.py:26-71)."""
    return dict(limit_angle=limit_angle, entity_attr=entity_attr, entity_manager=entity_manager,
                grace_steps=grace_steps)


@term("termination", "GFB_T_BASE_HEIGHT_MIN")
def base_height_below_minimum(env, minimum_height=0.05, entity_attr="robot", entity_manager=None):
    """Base z below a minimum (terminations.py:74-99)."""
    return dict(minimum_height=minimum_height, entity_attr=entity_attr, entity_manager=entity_manager)


@term("termination", "GFB_T_OUT_OF_BOUNDS")
def out_of_bounds(env, terrain_manager, subterrain=None, border_margin=0.5, entity_attr="robot"):
    """Base xy outside the (sub)terrain bounds shrunk by a margin (terminations.py:102-137)."""
    return dict(terrain_manager=terrain_manager, subterrain=subterrain, border_margin=border_margin,
                entity_attr=entity_attr)


@term("termination", "GFB_T_HAS_CONTACT")
def has_contact(_env, contact_manager, threshold=1.0, min_contacts=1):
    """At least `min_contacts` tracked links exceed the force threshold (terminations.py:139-155)."""
    return dict(contact_manager=contact_manager, threshold=threshold, min_contacts=min_contacts)


@term("termination", "GFB_T_CONTACT_FORCE")
def contact_force(_env, contact_manager, threshold=1.0):
    """Any tracked link exceeds the force threshold (terminations.py:158-172)."""
    return dict(contact_manager=contact_manager, threshold=threshold)


@term("termination", "GFB_T_CONTACT_FORCE_GRACE")
def contact_force_with_grace_period(env, contact_manager, threshold=100.0, grace_steps=10):
    """contact_force, ignored during the first `grace_steps` of an episode (terminations.py:175-205)."""
    return dict(contact_manager=contact_manager, threshold=threshold, grace_steps=grace_steps)
