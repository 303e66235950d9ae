"""Wrapper glue of the reference (genesis_forge/wrappers): the base Wrapper and the rsl_rl adapter."""
from .rsl_rl import RslRlWrapper
from .wrapper import Wrapper

__all__ = ["Wrapper", "RslRlWrapper"]
