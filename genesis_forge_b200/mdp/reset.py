"""
EntityManager on_reset items, by the reference's names (genesis_forge/mdp/reset.py).  These end in
engine setters (`set_pos`, `set_quat`, ...) for the compacted list of reset envs that the fused
kernel produced; they are host-side consumers of the hot path, not part of it.
"""
from __future__ import annotations

import math
from typing import Callable

import torch

from .._gs import gs
from ..managers.config import ResetMdpFnClass
from ..utils import links_by_name_pattern, xyz_to_quat


def zero_all_dofs_velocity(env, entity, envs_idx):
    entity.zero_all_dofs_velocity(envs_idx)


def set_rotation(env, entity, envs_idx, x=0, y=0, z=0):
    """Set (or uniformly randomise, when a (min, max) tuple is given) the entity's Euler rotation."""
    angles = torch.zeros((len(envs_idx), 3), device=gs.device)
    for col, (axis, value) in enumerate((("x", x), ("y", y), ("z", z))):
        if isinstance(value, tuple):
            angles[:, col] = env.rng.uniform(f"set_rotation_{axis}", angles[:, col], *value)
    entity.set_quat(xyz_to_quat(angles), envs_idx=envs_idx)


class position(ResetMdpFnClass):
    """Reset to a fixed position and (optional) rotation (reset.py:67-124)."""

    def __init__(self, env, entity, position, quat=None, zero_velocity: bool = True):
        self.env = env
        self.zero_velocity = zero_velocity
        n = env.num_envs
        self.reset_pos = torch.tensor(position, device=gs.device, dtype=gs.tc_float)
        # constant rows, filled once: a reset hands the engine the first n rows (a contiguous view)
        self._pos_rows = self.reset_pos.unsqueeze(0).repeat(n, 1)
        self.reset_quat = None
        self._quat_rows = None
        if quat is not None:
            self.reset_quat = torch.tensor(quat, device=gs.device, dtype=gs.tc_float)
            self._quat_rows = self.reset_quat.unsqueeze(0).repeat(n, 1)

    def __call__(self, env, entity, envs_idx, position, quat=None, zero_velocity: bool = True):
        n = len(envs_idx)
        entity.set_pos(self._pos_rows[:n], envs_idx=envs_idx, zero_velocity=self.zero_velocity)
        if self._quat_rows is not None:
            entity.set_quat(self._quat_rows[:n], envs_idx=envs_idx, zero_velocity=self.zero_velocity)


class randomize_terrain_position(ResetMdpFnClass):
    """
    Random spot on the terrain with a random yaw by default (reset.py:127-226).  Position, Euler
    draws and quaternion of all reset envs come from ONE launch of the library's spawn kernel
    (TerrainManager._spawn -> gfb_spawn_pose); the engine setters follow.
    """

    def __init__(self, env, entity, terrain_manager, height_offset: float = 0.1e-3,
                 subterrain: str | Callable[[], str] | None = None,
                 rotation: dict | None = {"z": (0, 2 * math.pi)}, zero_velocity: bool = True):
        self.env = env
        self.rotation = rotation
        self._rotation_buffer = None
        self._quat_buffer = None

    def build(self):
        self._rotation_buffer = torch.zeros((self.env.num_envs, 3), device=gs.device, dtype=gs.tc_float)
        self._quat_buffer = torch.zeros((self.env.num_envs, 4), device=gs.device, dtype=gs.tc_float)

    def __call__(self, env, entity, envs_idx, terrain_manager, height_offset: float = 0.1e-3, subterrain=None,
                 rotation: dict | None = {"z": (0, 2 * math.pi)}, zero_velocity: bool = True):
        if subterrain is not None and callable(subterrain):
            subterrain = subterrain()
        pos, quat = terrain_manager._spawn(
            terrain_manager._env_pos_buffer, envs_idx, 0.5, subterrain, height_offset, compact=True,
            rotation=rotation, rot_buffer=self._rotation_buffer, quat_buffer=self._quat_buffer,
        )
        entity.set_pos(pos, envs_idx=envs_idx, zero_velocity=zero_velocity)
        if rotation is not None:
            entity.set_quat(quat, envs_idx=envs_idx, zero_velocity=zero_velocity)


class randomize_link_mass_shift(ResetMdpFnClass):
    """Add a random mass shift to the links matching `link_name` on every reset (reset.py:229-284)."""

    def __init__(self, _env, entity, link_name: str, add_mass_range=(-0.2, 0.2)):
        self.env = _env
        self.add_mass_range = add_mass_range
        self._entity = entity
        self._link_name = link_name
        self.build()

    def build(self):
        self._links_idx_local = []
        self._mass_shift_buffer = None
        if self._link_name is not None:
            links = links_by_name_pattern(self._entity, self._link_name)
            if links:
                self._links_idx_local = [link.idx_local for link in links]
                self._mass_shift_buffer = torch.zeros(
                    (self.env.num_envs, len(self._links_idx_local)), device=gs.device
                )

    def __call__(self, env, entity, envs_idx, link_name: str, add_mass_range=(-0.2, 0.2)):
        fused = getattr(env, "_fused", None)
        if fused is not None and not fused.dry_run and torch.is_tensor(envs_idx):
            # U(lo, hi) per reset env and link, scattered into the persistent buffer: one launch
            lo, hi = self.add_mass_range
            fused.reset_rows("uniform", "mass_shift", envs_idx, envs_idx.shape[0], self._mass_shift_buffer.shape[1], lo, hi,
                             scatter=self._mass_shift_buffer)
        else:
            like = self._mass_shift_buffer[envs_idx, :]
            self._mass_shift_buffer[envs_idx, :] = env.rng.uniform("mass_shift", like, *self.add_mass_range)
        self._entity.set_mass_shift(self._mass_shift_buffer, links_idx_local=self._links_idx_local, envs_idx=envs_idx)
