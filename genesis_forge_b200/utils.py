"""
Entity helpers with the reference's names (genesis_forge/utils.py).

`entity_lin_vel / entity_ang_vel / entity_projected_gravity` are the "uncached" body-frame
transforms (they invert the entity's CURRENT quaternion instead of the EntityManager cache).  Inside
reward / termination configs the fused kernel evaluates them (before reset the two paths see the
same quaternion); called directly they run the library's rotation kernel.
"""
from __future__ import annotations

import re

import torch

from ._gs import gs


def _fused_of(entity):
    env = getattr(getattr(entity, "_scene", None), "_env", None)
    if env is None or getattr(env, "_fused", None) is None:
        raise RuntimeError(
            "entity_* helpers need an entity that belongs to a built ManagedEnvironment "
            "(the computation runs in the CUDA library; there is no eager fallback)"
        )
    return env._fused


def entity_lin_vel(entity) -> torch.Tensor:
    return _fused_of(entity).rotate_by_inv_quat(entity.get_vel(), entity.get_quat())


def entity_ang_vel(entity) -> torch.Tensor:
    return _fused_of(entity).rotate_by_inv_quat(entity.get_ang(), entity.get_quat())


def entity_projected_gravity(entity) -> torch.Tensor:
    return _fused_of(entity).rotate_by_inv_quat(None, entity.get_quat())


def links_by_name_pattern(entity, name_pattern: str) -> list:
    return [
        link for link in entity.links
        if link.name == name_pattern or re.match(f"^{name_pattern}$", link.name)
    ]


def xyz_to_quat(xyz: torch.Tensor) -> torch.Tensor:
    """Extrinsic x-y-z Euler angles (..., 3) -> (w, x, y, z).  Uses Genesis' own helper when present."""
    try:  # pragma: no cover
        from genesis.utils.geom import xyz_to_quat as _gs_xyz_to_quat  # type: ignore

        return _gs_xyz_to_quat(xyz)
    except Exception:
        pass
    half = xyz * 0.5
    cx, cy, cz = torch.cos(half[..., 0]), torch.cos(half[..., 1]), torch.cos(half[..., 2])
    sx, sy, sz = torch.sin(half[..., 0]), torch.sin(half[..., 1]), torch.sin(half[..., 2])
    return torch.stack(
        [
            cx * cy * cz - sx * sy * sz,
            sx * cy * cz + cx * sy * sz,
            cx * sy * cz - sx * cy * sz,
            cx * cy * sz + sx * sy * cz,
        ],
        dim=-1,
    )
