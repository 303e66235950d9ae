"""
GenesisEnv: base environment with the reference's public surface
(genesis_forge/genesis_env.py:12-286): counters, action buffers, per-env max episode length,
the `extras` dict rebuilt every step.

The step/reset bookkeeping the reference does here with eager ops (episode_length += 1, the
actions / last_actions ring, zeroing at reset, max_episode_length re-randomisation) is executed by
the fused kernels when the subclass is a ManagedEnvironment; this class keeps the state tensors
and the non-managed behaviour.
"""
from __future__ import annotations

import math
from typing import Any, Literal

import torch

from ._gs import gs
from .rng import HostRng

EnvMode = Literal["train", "eval", "play"]


class GenesisEnv:
    action_space = None
    observation_space = None
    can_be_wrapped: bool = True

    def __init__(
        self,
        num_envs: int = 1,
        dt: float = 1 / 100,
        max_episode_length_sec: int | None = 10,
        max_episode_random_scaling: float = 0.0,
        extras_logging_key: str = "episode",
    ):
        self.dt = dt
        self.device = gs.device
        self.num_envs = num_envs
        self.scene = None
        self.robot = None
        self.terrain = None
        self.rng = HostRng()

        self.extras_logging_key = extras_logging_key
        self._extras = {extras_logging_key: {}}
        self._actions: torch.Tensor | None = None
        self._last_actions: torch.Tensor | None = None

        self.step_count: int = 0
        self.episode_length = torch.zeros((num_envs,), device=gs.device, dtype=torch.int32)
        self.max_episode_length: torch.Tensor | None = None
        self._max_episode_length_sec = 0.0
        self._base_max_episode_length = None
        self._max_episode_random_scaling = max_episode_random_scaling
        if max_episode_length_sec and max_episode_length_sec > 0:
            self.max_episode_length = torch.zeros((num_envs,), device=gs.device, dtype=gs.tc_int)
            self.max_episode_length[:] = self.set_max_episode_length(max_episode_length_sec)

    # -- properties ---------------------------------------------------------------------------
    @property
    def unwrapped(self):
        return self

    @property
    def max_episode_length_sec(self):
        return self._max_episode_length_sec

    @property
    def extras(self) -> dict:
        return self._extras

    @property
    def actions(self) -> torch.Tensor:
        """Raw actions of this step (before the action manager)."""
        return self._actions

    @property
    def last_actions(self) -> torch.Tensor:
        return self._last_actions

    @property
    def num_actions(self) -> int:
        return self.action_space.shape[0] if self.action_space is not None else 0

    @property
    def num_observations(self) -> int:
        return self.observation_space.shape[0] if self.observation_space is not None else 0

    @property
    def max_episode_length_steps(self):
        return self._base_max_episode_length

    def set_max_episode_length(self, max_episode_length_sec) -> int:
        self._max_episode_length_sec = max_episode_length_sec
        self._base_max_episode_length = math.ceil(max_episode_length_sec / self.dt)
        return self._base_max_episode_length

    # -- operations ---------------------------------------------------------------------------
    def build(self) -> None:
        assert self.scene is not None, (
            "The scene must be constructed and assigned to the <env>.scene attribute before building."
        )
        self.scene.build(n_envs=self.num_envs)

    def _begin_step(self):
        """Fresh extras + step counter (genesis_env.py:193-195)."""
        self._extras = {self.extras_logging_key: {}}
        self.step_count += 1

    def _allocate_action_buffers(self, width: int):
        self._actions = torch.zeros((self.num_envs, width), device=gs.device, dtype=gs.tc_float)
        self._last_actions = torch.zeros_like(self._actions)

    def step(self, actions: torch.Tensor):
        """Non-managed bookkeeping step (a ManagedEnvironment overrides this with the fused path)."""
        self._begin_step()
        self.episode_length += 1
        if self._actions is None:
            self._actions = actions.detach().clone()
            self._last_actions = torch.zeros_like(actions, device=gs.device)
        else:
            self._last_actions.copy_(self._actions)
            self._actions.copy_(actions)
        return None, None, None, None, self._extras

    def reset(self, envs_idx=None) -> tuple[torch.Tensor | None, dict[str, Any]]:
        """Non-managed reset of the env-level buffers (genesis_env.py:207-254)."""
        if envs_idx is None:
            envs_idx = torch.arange(self.num_envs, device=gs.device)
        if self.step_count == 0 and self.action_space is not None and self._actions is None:
            self._allocate_action_buffers(self.action_space.shape[0])
        if envs_idx.numel() > 0:
            if self._actions is not None:
                self._actions[envs_idx] = 0.0
                self._last_actions[envs_idx] = 0.0
            self.episode_length[envs_idx] = 0
        if (
            len(envs_idx) > 0
            and self._max_episode_random_scaling > 0.0
            and self._base_max_episode_length is not None
        ):
            span = self._base_max_episode_length * self._max_episode_random_scaling
            u = self.rng.uniform("max_len", torch.empty((envs_idx.numel(),)), -1.0, 1.0)
            lengths = torch.round(self._base_max_episode_length + u * span).to(gs.tc_int)
            self.max_episode_length[envs_idx] = lengths.to(gs.device)
        return None, self.extras

    def get_observations(self) -> torch.Tensor | None:
        if self.observation_space is not None:
            return torch.zeros(
                (self.num_envs, self.observation_space.shape[0]), device=gs.device, dtype=gs.tc_float
            )
        return None

    def close(self):
        pass
