"""
Declarative reward / termination terms.

In the reference a term is a Python function `fn(env, **params) -> Tensor` that the manager calls
every step (reward_manager.py:185, termination_manager.py:168).  Here the functions with the
reference's names and signatures are *descriptors*: a manager config refers to them exactly as
before (`"fn": rewards.lin_vel_z_l2, "params": {...}`), the fused step recognises them by identity
and evaluates the corresponding opcode inside the kernel.  Calling one directly evaluates that one
term through the same kernel code path (no eager re-implementation exists).
"""
from __future__ import annotations

import functools


def term(kind: str, opcode: str):
    """Mark a function as a kernel term: kind 'reward' | 'termination', opcode = GFB_R_* / GFB_T_* name."""

    def wrap(fn):
        @functools.wraps(fn)
        def call(env, *args, **kwargs):
            fused = getattr(env, "_fused", None)
            if fused is None:
                raise RuntimeError(f"{fn.__name__}: the environment is not built yet")
            bound = fn(env, *args, **kwargs)  # validates arguments, returns the param dict
            return fused.evaluate_single_term(kind, call, bound)

        call.gfb_kind = kind
        call.gfb_opcode = opcode
        call.gfb_signature = fn
        return call

    return wrap
