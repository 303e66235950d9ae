"""
TEST INFRASTRUCTURE (oracle).  Runs the UNMODIFIED reference managers (imported from
/root/reference under oracle/shim.py) against the synthetic engine.

Exists only in this container (the GPU box has no /root/reference).  Used to
  * pin oracle/manager_port.py: the port must reproduce the reference bit for bit, step by step,
    on every config (tests/test_oracle_vs_reference.py, skipped when the reference is absent);
  * generate the committed golden traces under tests/golden/ (oracle/make_golden.py).

The one substitution made inside the reference is its Taichi contact kernel
(managers/contact/kernel.py), which cannot run without gstaichi: the symbol
`genesis_forge.managers.contact.contact_manager.kernel_get_contact_forces` is rebound to the ordered
torch restatement in oracle/contact_kernel.py.  Everything else executes the reference's own code.
"""
from __future__ import annotations

import os

import torch

from . import shim
from .contact_kernel import kernel_get_contact_forces
from configs.env_builder import build_env, reference_namespace


def reference_available() -> bool:
    return os.path.isdir(os.path.join(shim.REFERENCE_ROOT, "genesis_forge"))


def install_contact_kernel():
    import genesis_forge.managers.contact.contact_manager as cm

    cm.kernel_get_contact_forces = kernel_get_contact_forces


def make_reference_env(spec: dict, num_envs: int, source=None, n_contacts: int = 8, seed: int = 1234, **scene_kw):
    """Reference ManagedEnvironment (torch-CPU) for `spec`, not yet built."""
    shim.install("cpu")
    ns = reference_namespace()
    install_contact_kernel()
    env = build_env(
        spec, ns, num_envs, torch.device("cpu"),
        source=source, copy_on_get=True, n_contacts=n_contacts, seed=seed, **scene_kw,
    )
    return env
