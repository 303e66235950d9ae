"""
TEST INFRASTRUCTURE (oracle).  Snapshot / comparison helpers shared by the parity tests.

`reference_snapshot(env)` reads every persistent buffer of a reference-API environment (the
unmodified reference OR the drop-in, which exposes the same public attributes) into the flat key
space that `PortEnv.snapshot()` uses, so that any two of {reference, port, drop-in} can be compared.
"""
from __future__ import annotations

import torch


def reference_snapshot(env) -> dict[str, torch.Tensor]:
    s = {
        "episode_length": env.episode_length,
        "max_episode_length": env.max_episode_length,
        "actions": env.actions,
        "last_actions": env.last_actions,
        "targets": env.action_manager.get_actions(),
        "base_pos": env.robot_manager.base_pos,
        "base_quat": env.robot_manager.base_quat,
        "inv_base_quat": env.robot_manager.inv_base_quat,
        "terminated": env.termination_manager.terminated,
        "truncated": env.termination_manager.truncated,
        "reward_buf": env.reward_manager.rewards,
        "episode_seconds": env.reward_manager._episode_seconds,
    }
    for name, v in env.reward_manager.episode_data.items():
        s[f"episode_data/{name}"] = v
    for name in env.managers_by_name("command"):
        mgr = getattr(env, name)
        s[f"command/{name}"] = mgr._command
        if hasattr(mgr, "foot_offset"):  # the gait_trainer example's manager: its own state tensors
            for key in ("foot_offset", "gait_period", "foot_height", "gait_time", "gait_phase", "clock_input"):
                s[f"gait/{name}/{key}"] = getattr(mgr, key)
            s[f"gait/{name}/gait_selected"] = mgr._gait_selected
    for name in env.managers_by_name("contact"):
        m = getattr(env, name)
        s[f"contact/{name}/contacts"] = m.contacts
        s[f"contact/{name}/positions"] = m.contact_positions
        if m.last_air_time is not None:
            s[f"contact/{name}/last_air"] = m.last_air_time
            s[f"contact/{name}/cur_air"] = m.current_air_time
            s[f"contact/{name}/last_contact"] = m.last_contact_time
            s[f"contact/{name}/cur_contact"] = m.current_contact_time
    return {k: v.detach().cpu().clone() for k, v in s.items() if v is not None}


def diff_exact(a: dict, b: dict) -> list[str]:
    """Keys whose tensors are not bitwise identical (NaNs compare equal)."""
    bad = []
    for k in sorted(set(a) | set(b)):
        if k not in a or k not in b:
            bad.append(f"{k}: missing on one side")
            continue
        x, y = a[k], b[k]
        if x.shape != y.shape or x.dtype != y.dtype:
            bad.append(f"{k}: shape/dtype {tuple(x.shape)}/{x.dtype} vs {tuple(y.shape)}/{y.dtype}")
            continue
        same = torch.equal(x, y) if not x.is_floating_point() else bool(((x == y) | (x.isnan() & y.isnan())).all())
        if not same:
            n = int((~((x == y) | (x.isnan() & y.isnan()))).sum()) if x.is_floating_point() else int((x != y).sum())
            bad.append(f"{k}: {n} of {x.numel()} elements differ")
    return bad


def extras_to_cpu(extras: dict) -> dict[str, torch.Tensor]:
    """Flatten the logging part of `extras` ("episode" dict) to CPU scalars."""
    out = {}
    for k, v in extras.get("episode", {}).items():
        out[k] = torch.as_tensor(v).detach().cpu().reshape(()).clone()
    return out
