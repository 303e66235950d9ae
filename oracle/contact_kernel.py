"""
TEST INFRASTRUCTURE (oracle).  Ordered torch restatement of the reference's only device kernel,
`kernel_get_contact_forces` (genesis_forge/managers/contact/kernel.py:5-90, a Taichi @ti.kernel).

The reference accumulates with float atomics inside a parallel Taichi loop, so its summation order
is nondeterministic.  The oracle (and the CUDA kernel) fix the order: contact slots are visited in
index order c = 0..C-1 and each target link's accumulators receive one rounded add per slot.
"Bit-exact contact sums" in the parity tests means bit-exact against THIS ordering.

Same call signature as the reference kernel (contact_manager.py:414-426): the caller allocates and
zero-fills the three outputs; all tensors contiguous; returns nothing.
"""
from __future__ import annotations

import torch

from .geom import ti_inv_transform_by_quat


def kernel_get_contact_forces(
    contact_forces: torch.Tensor,    # (N, C, 3)
    contact_positions: torch.Tensor,  # (N, C, 3)
    link_a: torch.Tensor,            # (N, C) int
    link_b: torch.Tensor,            # (N, C) int
    links_quat: torch.Tensor,        # (N, L, 4)
    target_link_ids: torch.Tensor,   # (Lc,)
    with_link_ids: torch.Tensor,     # (Lw,)
    output_forces: torch.Tensor,     # (N, Lc, 3)  zero-filled
    output_positions: torch.Tensor,  # (N, Lc, 3)  zero-filled
    position_counts: torch.Tensor,   # (N, Lc)     zero-filled
    has_with_filter: int,
) -> None:
    n_envs, n_slots = link_a.shape
    rows = torch.arange(n_envs, device=link_a.device)
    la, lb = link_a.long(), link_b.long()
    targets = [int(t) for t in target_link_ids.tolist()]
    withs = [int(w) for w in with_link_ids.tolist()]
    zero3 = torch.zeros((), dtype=contact_forces.dtype, device=contact_forces.device)

    for c in range(n_slots):  # kernel.py:35-37, serialised over the contact slot index
        a, b = la[:, c], lb[:, c]
        force = contact_forces[:, c, :]
        pos = contact_positions[:, c, :]
        # kernel.py:68-78: the force is rotated into the TARGET link's frame; when the target is
        # link_b the force is taken as is, when it is link_a the reaction (-f) is used.
        f_if_b = ti_inv_transform_by_quat(force, links_quat[rows, b])
        f_if_a = ti_inv_transform_by_quat(-force, links_quat[rows, a])
        for t, target in enumerate(targets):
            is_a = a == target  # kernel.py:43-44
            is_b = b == target
            hit = is_a | is_b
            if has_with_filter:  # kernel.py:48-57
                keep = torch.zeros_like(hit)
                for w in withs:
                    keep |= (is_a & (b == w)) | (is_b & (a == w))
                hit = hit & keep
            local = torch.where(is_b.unsqueeze(-1), f_if_b, f_if_a)  # kernel.py:75 (b wins)
            mask = hit.unsqueeze(-1)
            output_positions[:, t, :] += torch.where(mask, pos, zero3)  # kernel.py:64
            position_counts[:, t] += hit.to(position_counts.dtype)     # kernel.py:65
            output_forces[:, t, :] += torch.where(mask, local, zero3)   # kernel.py:81-82

    # kernel.py:85-90: mean contact position per target link
    nonzero = position_counts > 0
    mean = output_positions / position_counts.clamp(min=1.0).unsqueeze(-1)
    output_positions.copy_(torch.where(nonzero.unsqueeze(-1), mean, output_positions))
