"""
Identities that hold for ANY correct implementation of the third-party quaternion helpers restated
in oracle/geom.py (genesis.utils.geom; parity unpinned -- the reference has no tests for them).
"""
import math

import torch

from oracle.geom import inv_quat, ti_inv_transform_by_quat, transform_by_quat, xyz_to_quat


def _unit_quats(n, seed=0):
    g = torch.Generator().manual_seed(seed)
    q = torch.randn(n, 4, generator=g)
    return q / q.norm(dim=1, keepdim=True)


def test_upright_gravity_and_norm_preservation():
    upright = torch.tensor([[1.0, 0.0, 0.0, 0.0]])
    g = torch.tensor([[0.0, 0.0, -1.0]])
    assert torch.equal(transform_by_quat(g, inv_quat(upright)), g)
    q = _unit_quats(1000)
    v = torch.randn(1000, 3, generator=torch.Generator().manual_seed(1))
    assert torch.allclose(transform_by_quat(v, q).norm(dim=1), v.norm(dim=1), rtol=1e-5, atol=1e-6)


def test_inverse_round_trip_and_involution():
    q = _unit_quats(1000, 2)
    v = torch.randn(1000, 3, generator=torch.Generator().manual_seed(3))
    assert torch.equal(inv_quat(inv_quat(q)), q)
    back = transform_by_quat(transform_by_quat(v, q), inv_quat(q))
    assert torch.allclose(back, v, atol=1e-5)


def test_taichi_form_agrees_with_torch_form():
    q = _unit_quats(1000, 4)
    v = torch.randn(1000, 3, generator=torch.Generator().manual_seed(5))
    a = ti_inv_transform_by_quat(v, q)
    b = transform_by_quat(v, inv_quat(q))
    assert torch.allclose(a, b, atol=1e-5)


def test_known_rotation_and_euler():
    # +90 deg about z maps x -> y
    s = math.sqrt(0.5)
    q = torch.tensor([[s, 0.0, 0.0, s]])
    out = transform_by_quat(torch.tensor([[1.0, 0.0, 0.0]]), q)
    assert torch.allclose(out, torch.tensor([[0.0, 1.0, 0.0]]), atol=1e-6)
    assert torch.allclose(xyz_to_quat(torch.tensor([[0.0, 0.0, math.pi / 2]])), q, atol=1e-6)
    assert torch.allclose(xyz_to_quat(torch.zeros(1, 3)), torch.tensor([[1.0, 0.0, 0.0, 0.0]]))
