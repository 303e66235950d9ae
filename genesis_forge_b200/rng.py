"""
Host-side random draws of the manager path (reset-time domain randomisation, spawn positions).

The reference draws these with `tensor.uniform_()` / `torch.rand_like()` on `gs.device`
(position_action_manager.py:516-525, mdp/reset.py:181-191, terrain_manager.py:230-235).  Every such
draw in this package goes through one object so that a parity harness can replay the reference's
own draws (BASELINE.json: "random command and domain-randomisation draws injected from the
reference").  Draws made INSIDE the kernels (command resample, max episode length, observation
noise) are either Philox (production) or dense injected buffers (parity), see fused.py.
"""
from __future__ import annotations

import torch


class HostRng:
    """Default: torch's generator for the tensor's device."""

    def uniform(self, tag: str, like: torch.Tensor, lo: float, hi: float) -> torch.Tensor:
        return torch.empty_like(like).uniform_(lo, hi)

    def multinomial(self, tag: str, weights: torch.Tensor) -> torch.Tensor:
        """One category per row of `weights` (user-level managers: the gait_trainer example's gait choice)."""
        return torch.multinomial(weights, 1).squeeze(-1)


class ReplayRng(HostRng):
    """Returns recorded draws, in recording order per tag; raises when a tag runs dry."""

    def __init__(self):
        self.queues: dict[str, list[torch.Tensor]] = {}

    def push(self, tag: str, value: torch.Tensor):
        self.queues.setdefault(tag, []).append(value)

    def clear(self):
        self.queues.clear()

    def uniform(self, tag, like, lo, hi):
        q = self.queues.get(tag)
        if not q:
            raise RuntimeError(f"ReplayRng: no recorded draw left for '{tag}'")
        v = q.pop(0).to(like.device, like.dtype)
        if v.shape != like.shape:
            raise RuntimeError(f"ReplayRng: '{tag}' shape {tuple(v.shape)} != {tuple(like.shape)}")
        return v

    def multinomial(self, tag, weights):
        q = self.queues.get(tag)
        if not q:
            raise RuntimeError(f"ReplayRng: no recorded draw left for '{tag}'")
        v = q.pop(0).to(weights.device)
        if v.shape[0] != weights.shape[0]:
            raise RuntimeError(f"ReplayRng: '{tag}' has {v.shape[0]} rows, {weights.shape[0]} requested")
        return v
