"""
Manager base class.  Same protocol as the reference's BaseManager
(genesis_forge/managers/base.py:16-43): a manager registers itself with the environment under a
type name on construction and offers build() / step() / reset(envs_idx).

In this implementation the per-step arithmetic of every manager runs inside the fused CUDA step
that ManagedEnvironment drives (genesis_forge_b200/fused.py); the manager objects hold the
configuration, own the state tensors the kernels read and write, and expose the reference's public
attributes on top of them.
"""
from __future__ import annotations

from typing import Literal

ManagerType = Literal[
    "action", "reward", "termination", "contact", "terrain", "entity", "command", "observation",
]


class BaseManager:
    def __init__(self, env, type: ManagerType, enabled: bool = True):
        self.env = env
        # the reference ignores the `enabled` argument (base.py:28); kept for parity
        self.enabled = True
        self.type = type
        if hasattr(env, "add_manager"):
            env.add_manager(type, self)

    def build(self):
        """Called when the scene is built."""

    def step(self):
        """Called when the environment is stepped (a no-op: the fused step does the work)."""

    def reset(self, envs_idx=None):
        """One or more environments have been reset."""
