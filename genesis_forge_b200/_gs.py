"""
Engine namespace used by the host classes (`gs.device`, `gs.tc_float`, ...).

The reference reads these from the `genesis` module (e.g. genesis_forge/genesis_env.py:62,76,85).
When Genesis is importable the real module's values are used; otherwise (benchmarks on synthetic
physics state, tests) a small namespace provides the same names.  `set_device()` selects the device
the managers allocate on in the synthetic case.
"""
from __future__ import annotations

import enum

import torch


class _JointType(enum.Enum):
    FIXED = 0
    REVOLUTE = 1
    PRISMATIC = 2
    FREE = 3


class _Namespace:
    """Mirror of the handful of `genesis` module attributes the manager path touches."""

    def __init__(self):
        self._real = None
        self.device = torch.device("cuda", 0) if torch.cuda.is_available() else torch.device("cpu")
        self.tc_float = torch.float32
        self.tc_int = torch.int32
        self.tc_bool = torch.bool
        self.JOINT_TYPE = _JointType
        self.GenesisException = RuntimeError

    def bind_real(self, module) -> None:
        """Adopt the constants of an initialised Genesis module (`gs.init()` already called)."""
        self._real = module
        for name in ("device", "tc_float", "tc_int", "tc_bool", "JOINT_TYPE"):
            if hasattr(module, name):
                setattr(self, name, getattr(module, name))


gs = _Namespace()


def set_device(device) -> torch.device:
    gs.device = torch.device(device)
    return gs.device


def try_bind_genesis() -> bool:
    """Use the real Genesis constants when the package is installed and initialised."""
    try:
        import genesis as real  # type: ignore
    except Exception:
        return False
    if getattr(real, "device", None) is None:
        return False
    gs.bind_real(real)
    return True
