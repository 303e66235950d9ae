"""
TerminationManager: OR-reduction of termination terms into `terminated` / `truncated`.

API of genesis_forge/managers/termination_manager.py.  The term evaluation, the OR into the two
masks (:162-175) and the per-term fire fractions logged when a term fired (:178-182) come out of
the fused post-physics kernel; this object owns the mask tensors and the live config.
"""
from __future__ import annotations

import torch

from .._gs import gs
from .base import BaseManager
from .config import TerminationConfigItem


class TerminationManager(BaseManager):
    def __init__(self, env, term_cfg: dict[str, dict], logging_enabled: bool = True, logging_tag: str = "Terminations"):
        super().__init__(env, type="termination")
        self.logging_enabled = logging_enabled
        self.logging_tag = logging_tag
        self.term_cfg: dict[str, TerminationConfigItem] = {
            name: TerminationConfigItem(cfg, env) for name, cfg in term_cfg.items()
        }
        self._terminated_buf = torch.zeros(env.num_envs, device=gs.device, dtype=torch.bool)
        self._truncated_buf = torch.zeros_like(self._terminated_buf)

    @property
    def dones(self) -> torch.Tensor:
        return self._terminated_buf | self._truncated_buf

    @property
    def terminated(self) -> torch.Tensor:
        return self._terminated_buf

    @property
    def truncated(self) -> torch.Tensor:
        return self._truncated_buf

    def build(self):
        for cfg in self.term_cfg.values():
            cfg.build()
