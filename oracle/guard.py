"""
TEST INFRASTRUCTURE (oracle).  Guard band around float-derived thresholds.

Boolean masks must match the reference bit for bit, but a handful of them are thresholds on values
that went through arithmetic whose last ulp is not portable between a CPU and a GPU (asin, sqrt of
multi-contact sums).  The kernels follow the torch-CPU op order exactly, so in
practice the values agree to the bit; to keep the parity tests meaningful on ANY host CPU the
synthetic physics states are nudged so that no env sits within a relative band of a threshold.
Which envs were nudged is counted and reported by the tests; nothing is hidden.

`make_sanitizer(spec, robot, terrain)` returns a `post(state, step_index)` hook for
genesis_forge_b200.synthetic.CachedSource.
"""
from __future__ import annotations

import math
import re

import torch

from .contact_kernel import kernel_get_contact_forces
from .geom import inv_quat, transform_by_quat

BAND = 2e-5  # relative


def _link_ids(entity, names):
    ids = []
    if names is None:
        return [link.idx for link in entity.links]
    for pattern in names:
        for link in entity.links:
            if pattern == link.name or re.match(f"^{pattern}$", link.name):
                ids.append(link.idx)
    return ids


def thresholds_of(spec: dict) -> tuple[list[float], dict[str, list[float]]]:
    """(tilt-angle limits in radians, per contact manager force thresholds) used anywhere in the spec."""
    tilt = []
    contact: dict[str, list[float]] = {name: [] for name in spec["contacts"]}
    for name, c in spec["contacts"].items():
        if c.get("track_air_time"):
            contact[name].append(c.get("air_time_contact_threshold", 1.0))
    for item in spec["terminations"].values():
        p = item.get("params") or {}
        if item["fn"] == "bad_orientation":
            tilt.append(math.radians(p.get("limit_angle", 40.0)))
        if "contact_manager" in p:
            default = 100.0 if item["fn"] == "contact_force_with_grace_period" else 1.0
            contact[p["contact_manager"][1:]].append(p.get("threshold", default))
    for item in spec["rewards"].values():
        p = item.get("params") or {}
        if "contact_manager" in p:
            mgr = p["contact_manager"][1:]
            if item["fn"] in ("has_contact", "contact_force"):
                contact[mgr].append(p.get("threshold", 1.0))
            if item["fn"] == "feet_slide":
                contact[mgr].append(1.0)
    return tilt, contact


def make_sanitizer(spec: dict, robot, terrain, band: float = BAND):
    tilt_limits, contact_thr = thresholds_of(spec)
    managers = {}
    for name, c in spec["contacts"].items():
        entity = robot
        ids = _link_ids(entity, c["link_names"])
        withs, has_filter = [], c.get("with_entity_attr") is not None or c.get("with_links_names") is not None
        if has_filter:
            w_entity = terrain if c.get("with_entity_attr") == "terrain" else robot
            withs = _link_ids(w_entity, c.get("with_links_names"))
        managers[name] = (torch.tensor(ids), torch.tensor(withs, dtype=torch.int64), has_filter)
    stats = {"nudged_tilt": 0, "nudged_contact": 0, "states": 0}

    def near(value: torch.Tensor, thr: float) -> torch.Tensor:
        return (value - thr).abs() <= band * max(abs(thr), 1e-3)

    def post(state: dict, step_index: int) -> dict:
        stats["states"] += 1
        n = state["quat"].shape[0]
        gravity = torch.tensor([0.0, 0.0, -1.0]).repeat(n, 1)
        for _ in range(20):
            bad_tilt = torch.zeros(n, dtype=torch.bool)
            if tilt_limits:
                g = transform_by_quat(gravity, inv_quat(state["quat"]))
                angle = torch.asin(torch.clamp(torch.norm(g[:, :2], dim=1), max=0.99))
                for limit in tilt_limits:
                    bad_tilt |= near(angle, limit)
            bad_contact = torch.zeros(n, dtype=torch.bool)
            for name, (ids, withs, has_filter) in managers.items():
                if not contact_thr[name]:
                    continue
                lc = ids.shape[0]
                out_f, out_p, cnt = torch.zeros(n, lc, 3), torch.zeros(n, lc, 3), torch.zeros(n, lc)
                kernel_get_contact_forces(
                    state["c_force"], state["c_pos"], state["c_link_a"], state["c_link_b"], state["links_quat"],
                    ids, withs, out_f, out_p, cnt, 1 if has_filter else 0,
                )
                norm = torch.norm(out_f, dim=-1)
                for thr in contact_thr[name]:
                    bad_contact |= near(norm, thr).any(dim=1)
            if not bad_tilt.any() and not bad_contact.any():
                break
            if bad_tilt.any():
                stats["nudged_tilt"] += int(bad_tilt.sum())
                q = state["quat"]
                q[bad_tilt, 1:] *= 1.03
                state["quat"] = q / q.norm(dim=1, keepdim=True)
            if bad_contact.any():
                stats["nudged_contact"] += int(bad_contact.sum())
                state["c_force"][bad_contact] *= 1.03
        else:
            raise RuntimeError("guard band: could not move every env away from the thresholds")
        return state

    post.stats = stats
    return post
