"""
TEST INFRASTRUCTURE (oracle).  Stub third-party modules so that the UNMODIFIED reference package
(/root/reference/genesis_forge) imports and runs on torch-CPU in a container that has neither
Genesis nor gstaichi/gymnasium/tensordict/hid/skrl.  This is the same trick the reference's own
docs build uses (docs/conf.py:14-24 `autodoc_mock_imports`), except that the stubs here are
functional where the manager path touches them:

    genesis                      device / tc_float / tc_int / tc_bool / JOINT_TYPE / Scene
    genesis.utils.geom           -> oracle/geom.py (restated arithmetic, parity unpinned)
    gstaichi                     `@ti.kernel` = identity decorator (the kernel body itself is replaced
                                 by the ordered restatement, see ref_harness.install_contact_kernel)
    gymnasium.spaces.Box         4-field record
    tensordict.TensorDict        dict subclass accepting (source, device=, batch_size=)
    hid, skrl...base.Wrapper     empty

Only oracle/ref_harness.py and tests that pin the port against the reference use this module; it
is never active in the product path.
"""
from __future__ import annotations

import sys
import types

import torch

from genesis_forge_b200._gs import gs as _engine_gs

from . import geom as _geom

import os as _os

# the reference itself (build container), else the verbatim copy made by oracle/make_ref.py (GPU box)
_COPY = _os.path.join(_os.path.dirname(_os.path.abspath(__file__)), "_ref")
REFERENCE_ROOT = "/root/reference" if _os.path.isdir("/root/reference/genesis_forge") else _COPY
_INSTALLED = False


class _Box:
    """Stand-in for gymnasium.spaces.Box (only the fields the reference reads)."""

    def __init__(self, low, high, shape=None, dtype=None):
        self.low, self.high, self.shape, self.dtype = low, high, tuple(shape), dtype

    def __repr__(self):
        return f"Box(shape={self.shape})"


class _TensorDict(dict):
    """Stand-in for tensordict.TensorDict: a dict that swallows device/batch_size."""

    def __init__(self, source=None, device=None, batch_size=None, **_):
        super().__init__(source or {})
        self.device = device
        self.batch_size = batch_size


def _module(name: str, **attrs) -> types.ModuleType:
    mod = types.ModuleType(name)
    mod.__dict__.update(attrs)
    sys.modules[name] = mod
    return mod


def install(device: str = "cpu") -> types.ModuleType:
    """Register the stub modules (idempotent) and return the fake `genesis` module."""
    global _INSTALLED
    if _INSTALLED:
        gs = sys.modules["genesis"]
        gs.device = torch.device(device)
        return gs

    class _Anything:
        def __init__(self, *a, **k):
            pass

    gs = _module(
        "genesis",
        device=torch.device(device),
        tc_float=torch.float32,
        tc_int=torch.int32,
        tc_bool=torch.bool,
        JOINT_TYPE=_engine_gs.JOINT_TYPE,  # the synthetic engine's joints carry this enum
        Scene=_Anything,
        GenesisException=RuntimeError,
    )
    engine = _module("genesis.engine")
    entities = _module("genesis.engine.entities", RigidEntity=_Anything)
    rigid_entity = _module("genesis.engine.entities.rigid_entity")
    rigid_link = _module("genesis.engine.entities.rigid_entity.rigid_link", RigidLink=_Anything)
    utils = _module("genesis.utils")
    geom = _module(
        "genesis.utils.geom",
        transform_by_quat=_geom.transform_by_quat,
        inv_quat=_geom.inv_quat,
        xyz_to_quat=_geom.xyz_to_quat,
        ti_inv_transform_by_quat=_geom.ti_inv_transform_by_quat,
    )
    vis = _module("genesis.vis")
    camera = _module("genesis.vis.camera", Camera=_Anything)
    gs.engine, engine.entities = engine, entities
    entities.rigid_entity, rigid_entity.rigid_link = rigid_entity, rigid_link
    gs.utils, utils.geom = utils, geom
    gs.vis, vis.camera = vis, camera

    class _Types:
        @staticmethod
        def ndarray(*a, **k):
            return None

    _module("gstaichi", kernel=lambda f: f, func=lambda f: f, types=_Types, i32=int, f32=float)

    spaces = _module("gymnasium.spaces", Box=_Box, Space=_Box)
    _module("gymnasium", spaces=spaces)
    _module("tensordict", TensorDict=_TensorDict)
    _module("hid")

    class _SkrlWrapper:
        def __init__(self, *a, **k):
            pass

    for name in ("skrl", "skrl.envs", "skrl.envs.wrappers", "skrl.envs.wrappers.torch"):
        _module(name)
    _module("skrl.envs.wrappers.torch.base", Wrapper=_SkrlWrapper)

    _INSTALLED = True
    return gs


def import_reference():
    """Import the unmodified reference package from /root/reference under the stubs."""
    import os

    if not os.path.isdir(os.path.join(REFERENCE_ROOT, "genesis_forge")):
        raise FileNotFoundError(f"{REFERENCE_ROOT}/genesis_forge not present (GPU box?)")
    install()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import genesis_forge  # noqa: F401
    import genesis_forge.managers  # noqa: F401
    import genesis_forge.mdp  # noqa: F401

    return sys.modules["genesis_forge"]
