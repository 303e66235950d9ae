"""
EntityManager: cached base pose of one entity, body-frame vectors, and on_reset hooks.

API of genesis_forge/managers/entity_manager.py.  The per-step cache refresh (:189-195) is the
first phase of the fused post-physics kernel, which writes `base_pos`, `base_quat` and
`inv_base_quat` into the tensors owned here.  As in the reference the cache is NOT refreshed by
reset(), so observations of freshly reset envs use the pre-reset quaternion.
on_reset items run on the host for the compacted reset index list the kernel produced.

An environment may register several EntityManagers (managed_env.py:200-220).  The kernel serves the
FIRST one; every further manager refreshes its cache in step() with device copies and evaluates its
body-frame getters through the library's rotation entry point (gfb_rotate), so observation terms
around them run as host-evaluated columns.
"""
from __future__ import annotations

import torch

from .._gs import gs
from .base import BaseManager
from .config import ConfigItem


class EntityManager(BaseManager):
    def __init__(self, env, entity_attr: str, on_reset: dict[str, dict]):
        super().__init__(env, type="entity")
        if hasattr(env, "add_entity_manager"):
            env.add_entity_manager(self)
        self.entity = None
        self._entity_attr = entity_attr
        self.on_reset: dict[str, ConfigItem] = {name: ConfigItem(cfg, env) for name, cfg in on_reset.items()}
        n = env.num_envs
        self._base_pos = torch.zeros((n, 3), device=gs.device, dtype=gs.tc_float)
        self._base_quat = torch.zeros((n, 4), device=gs.device, dtype=gs.tc_float)
        self._inv_base_quat = torch.zeros_like(self._base_quat)
        self._conjugate = torch.tensor([1.0, -1.0, -1.0, -1.0], device=gs.device, dtype=gs.tc_float)

    @property
    def _primary(self) -> bool:
        """The environment's first EntityManager: the one whose cache the post-physics kernel writes."""
        registered = getattr(self.env, "managers", {}).get("entity")
        return not registered or registered[0] is self

    @property
    def base_pos(self) -> torch.Tensor:
        return self._base_pos

    @property
    def base_quat(self) -> torch.Tensor:
        return self._base_quat

    @property
    def inv_base_quat(self) -> torch.Tensor:
        return self._inv_base_quat

    # body-frame vectors (entity_manager.py:130-146).  Inside an observation config these resolve to
    # kernel column sources; called directly they are evaluated by the library's rotation kernel.
    def get_projected_gravity(self) -> torch.Tensor:
        if not self._primary:
            return self._rotated("entity_gravity_b", None)
        return self.env._trace_or("gravity_b", lambda: self.env._fused.rotate_by_inv_base_quat(None))

    def get_linear_velocity(self) -> torch.Tensor:
        if not self._primary:
            return self._rotated("entity_lin_vel_b", self.entity.get_vel)
        return self.env._trace_or("lin_vel_b", lambda: self.env._fused.rotate_by_inv_base_quat(self.entity.get_vel()))

    def get_angular_velocity(self) -> torch.Tensor:
        if not self._primary:
            return self._rotated("entity_ang_vel_b", self.entity.get_ang)
        return self.env._trace_or("ang_vel_b", lambda: self.env._fused.rotate_by_inv_base_quat(self.entity.get_ang()))

    def _rotated(self, tag: str, getter) -> torch.Tensor:
        """A further manager's body-frame vector: rotation by the inverse of ITS cached quaternion."""
        return self.env._trace_or(
            (tag, self),
            lambda: self.env._fused.rotate_by_inv_quat(getter() if getter is not None else None, self._base_quat),
        )

    def build(self):
        self.entity = getattr(self.env, self._entity_attr)
        if not self._primary:
            self._cached_calcs()  # (the first manager's cache is filled by FusedStep.cache_entity)
        for cfg in self.on_reset.values():
            cfg.build(entity=self.entity)

    def step(self):
        """Per-step cache refresh of a further manager (entity_manager.py:163-167); the first manager's
        is the entity phase of the post-physics kernel."""
        if not self._primary:
            self._cached_calcs()

    def _cached_calcs(self):
        """entity_manager.py:189-195 with device copies; inv_quat is the conjugate (sign flips are exact)."""
        self._base_pos.copy_(self.entity.get_pos())
        self._base_quat.copy_(self.entity.get_quat())
        torch.mul(self._base_quat, self._conjugate, out=self._inv_base_quat)

    def reset(self, envs_idx=None):
        """Run the on_reset items for `envs_idx` (entity_manager.py:169-183)."""
        if not self.enabled:
            return
        if envs_idx is None:
            envs_idx = torch.arange(self.env.num_envs, device=gs.device)
        for name, cfg in self.on_reset.items():
            try:
                cfg.execute(envs_idx)
            except Exception:
                print(f"Error resetting entity with config: '{name}'")
                raise
