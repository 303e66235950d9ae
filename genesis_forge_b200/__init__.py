"""
genesis_forge_b200: B200-native implementation of genesis-forge's per-step manager pipeline.

Same manager classes, config dicts and ManagedEnvironment API as jgillick/genesis-forge 0.2.1; the
arithmetic of a step runs in hand-written sm_100a kernels behind the C ABI in include/gfb200.h.
"""
from .genesis_env import GenesisEnv, EnvMode
from .managed_env import ManagedEnvironment
from ._gs import gs, set_device

__all__ = ["GenesisEnv", "ManagedEnvironment", "EnvMode", "gs", "set_device"]
__version__ = "0.1.0"
