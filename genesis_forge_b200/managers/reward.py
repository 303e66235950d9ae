"""
RewardManager: weighted sum of reward terms, per-term episode sums, episode-mean logging at reset.

API of genesis_forge/managers/reward_manager.py.  Evaluation of every term, the `* weight * dt`
scaling, the running sums (:176-193) and the reset-time per-term means (:197-222) are computed in
the fused post-physics kernel; `cfg[name].weight` and `cfg[name].params` are read again on every
step, so curriculum code that mutates them keeps working.
"""
from __future__ import annotations

import torch

from .._gs import gs
from .base import BaseManager
from .config import RewardConfigItem


class RewardManager(BaseManager):
    def __init__(self, env, cfg: dict[str, dict], logging_enabled: bool = True, logging_tag: str = "Rewards"):
        super().__init__(env, type="reward")
        self.logging_enabled = logging_enabled
        self.logging_tag = logging_tag
        self.cfg: dict[str, RewardConfigItem] = {name: RewardConfigItem(c, env) for name, c in cfg.items()}
        n = env.num_envs
        self._reward_buf = torch.zeros((n,), device=gs.device, dtype=gs.tc_float)
        self._episode_seconds = torch.zeros((n,), device=gs.device, dtype=gs.tc_float)
        # one contiguous (T_r, N) block; episode_data[name] is row i (coalesced per-term access)
        self._episode_sums = torch.zeros((max(len(self.cfg), 1), n), device=gs.device, dtype=gs.tc_float)
        self._episode_data = {name: self._episode_sums[i] for i, name in enumerate(self.cfg)}
        self._episode_mean: dict[str, float] = {}
        self._last_log = None  # (device vector, {name: index}) of the most recent reset

    @property
    def rewards(self) -> torch.Tensor:
        return self._reward_buf

    @property
    def episode_data(self) -> dict[str, torch.Tensor]:
        return self._episode_data

    def last_episode_mean_reward(self, name: str, before_weight: bool = True) -> float:
        """Mean of the last finished episodes for one term (reward_manager.py:138-153)."""
        if self._last_log is not None:
            vec, index = self._last_log
            host = vec.tolist()
            for key, i in index.items():
                self._episode_mean[key] = host[i]
            self._last_log = None
        rew = self._episode_mean.get(name, 0.0)
        if before_weight:
            rew /= self.cfg[name].weight
        return rew

    def build(self):
        for cfg in self.cfg.values():
            cfg.build()
