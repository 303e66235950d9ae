"""
TEST INFRASTRUCTURE (oracle).  Restated quaternion arithmetic of `genesis.utils.geom`.

PARITY UNPINNED: these functions belong to the third-party package genesis-world (constraint
`genesis-world>=0.3.4` in the reference's pyproject.toml:12; no lock file, no vendored copy, source
absent from /root/reference and from this container).  The reference has no tests or golden vectors
for them.  What is restated here is the published algorithm (rotate v by unit quaternion q, w-first):

    transform_by_quat(v, q):   t = 2 * (q_xyz x v);  v' = v + w * t + q_xyz x t
    inv_quat(q):               (w, -x, -y, -z)
    ti_inv_transform_by_quat:  u = q*_xyz x v;  uu = q*_xyz x u;  v' = v + 2 * (w * u + uu),  q* = conj(q)
    xyz_to_quat(euler_xyz):    extrinsic x-y-z Euler angles -> (w, x, y, z)

Reference call sites (the only things that constrain these functions):
    genesis_forge/utils.py:23-24,37-38,51-55      entity_lin_vel / entity_ang_vel / entity_projected_gravity
    genesis_forge/managers/entity_manager.py:134,140,146,195
    genesis_forge/managers/contact/kernel.py:76,78  (ti_inv_transform_by_quat)
    genesis_forge/mdp/reset.py:63,194               (xyz_to_quat, reset side)

The CUDA kernels in genesis_forge_b200/csrc follow THIS file op for op (same association order,
every multiply/add separately rounded, cross products as fma(a1,b2,-(a2*b1)) which is what
torch.cross evaluates to on the CPU dispatch levels with FMA), so that masks derived from these
values are bit-identical between the kernels and the torch-CPU oracle.  Identities that hold for any
correct implementation are asserted in tests/test_oracle_geom.py.
"""
from __future__ import annotations

import torch


def inv_quat(quat: torch.Tensor) -> torch.Tensor:
    """Conjugate of a (w, x, y, z) quaternion; one elementwise multiply by (1,-1,-1,-1)."""
    sign = torch.tensor([1.0, -1.0, -1.0, -1.0], dtype=quat.dtype, device=quat.device)
    return quat * sign


def transform_by_quat(v: torch.Tensor, quat: torch.Tensor) -> torch.Tensor:
    """Rotate v (..., 3) by unit quaternion quat (..., 4), w first."""
    qvec = quat[..., 1:]
    t = torch.cross(qvec, v, dim=-1) * 2
    return v + quat[..., :1] * t + torch.cross(qvec, t, dim=-1)


def ti_inv_transform_by_quat(v: torch.Tensor, quat: torch.Tensor) -> torch.Tensor:
    """
    Torch rendering of the Taichi helper used inside the reference's contact kernel
    (kernel.py:76,78): rotate v by the conjugate of quat.  Written in the association order the
    CUDA contact path follows.
    """
    w = quat[..., :1]
    qvec = -quat[..., 1:]
    u = torch.cross(qvec, v, dim=-1)
    uu = torch.cross(qvec, u, dim=-1)
    return v + (w * u + uu) * 2


def xyz_to_quat(xyz: torch.Tensor, rpy: bool = False, degrees: bool = False) -> torch.Tensor:
    """Extrinsic x-y-z Euler angles (..., 3) -> quaternion (..., 4), w first.  Reset side only."""
    if degrees:
        xyz = torch.deg2rad(xyz)
    half = xyz * 0.5
    cx, cy, cz = torch.cos(half[..., 0]), torch.cos(half[..., 1]), torch.cos(half[..., 2])
    sx, sy, sz = torch.sin(half[..., 0]), torch.sin(half[..., 1]), torch.sin(half[..., 2])
    if rpy:
        w = cx * cy * cz + sx * sy * sz
        x = sx * cy * cz - cx * sy * sz
        y = cx * sy * cz + sx * cy * sz
        z = cx * cy * sz - sx * sy * cz
    else:
        w = cx * cy * cz - sx * sy * sz
        x = sx * cy * cz + cx * sy * sz
        y = cx * sy * cz - sx * cy * sz
        z = cx * cy * sz + sx * sy * cz
    return torch.stack([w, x, y, z], dim=-1)
