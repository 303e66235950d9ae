"""
Action managers: raw policy actions -> joint position targets.

API of genesis_forge/managers/action/{base,position_action_manager,position_within_limits}.py.
The per-step arithmetic (scale, offset, clip, NaN/Inf detection; position_action_manager.py:389-419,
position_within_limits.py:113-131) runs in the pre-physics kernel launched by the fused step
(gfb_action_step); what stays here is the joint-pattern -> per-DOF parameter resolution at build
time (position_action_manager.py:297-374, 470-514) and the reset-time engine writes with optional
domain-randomisation noise (position_action_manager.py:421-464, 516-525).
"""
from __future__ import annotations

import re
from typing import Any, Callable, TypeVar

import numpy as np
import torch

from .._gs import gs
from ..spaces import Box
from .base import BaseManager

T = TypeVar("T")
DofValue = dict[str, T] | T


def _ensure_dof_pattern(value):
    """A scalar means "every joint": 50 -> {".*": 50}."""
    if value is None:
        return None
    if isinstance(value, dict):
        return value
    return {".*": value}


class BaseActionManager(BaseManager):
    def __init__(self, env, delay_step: int = 0):
        super().__init__(env, type="action")
        self._raw_actions = None
        self._actions = None
        self._delay_step = delay_step
        self._action_delay_buffer: list[torch.Tensor] = []

    @property
    def num_actions(self) -> int:
        return 0

    @property
    def action_space(self):
        return Box(low=-np.inf, high=np.inf, shape=(self.num_actions,), dtype=np.float32)

    @property
    def actions(self) -> torch.Tensor:
        if self._actions is None:
            return torch.zeros((self.env.num_envs, self.num_actions))
        return self._actions

    @property
    def raw_actions(self) -> torch.Tensor:
        if self._raw_actions is None:
            return torch.zeros((self.env.num_envs, self.num_actions))
        return self._raw_actions

    def _delayed(self, actions: torch.Tensor) -> torch.Tensor:
        """The reference's delay FIFO (base.py:72-74): newest in at the front, oldest out at the back."""
        if self._delay_step > 0:
            self._action_delay_buffer.insert(0, actions)
            actions = self._action_delay_buffer.pop()
        self._raw_actions = actions
        return actions

    def get_actions(self) -> torch.Tensor:
        if self._actions is None:
            return self.env._trace_or(
                "targets", lambda: torch.zeros((self.env.num_envs, self.num_actions), device=gs.device)
            )
        return self.env._trace_or("targets", lambda: self._actions)


class PositionActionManager(BaseActionManager):
    """position = offset + scale * action, clipped; offset defaults to the default joint pose."""

    kernel_mode = 1

    def __init__(
        self,
        env,
        joint_names: list[str] | str = ".*",
        default_pos: DofValue[float] = {".*": 0.0},
        scale: DofValue[float] = 1.0,
        clip: DofValue[tuple[float, float]] = None,
        offset: DofValue[float] = 0.0,
        use_default_offset: bool = True,
        pd_kp: DofValue[float] = None,
        pd_kv: DofValue[float] = None,
        max_force: DofValue[float | tuple[float, float]] = None,
        damping: DofValue[float] = None,
        stiffness: DofValue[float] = None,
        frictionloss: DofValue[float] = None,
        noise_scale: float = 0.0,
        action_handler: Callable[[torch.Tensor], None] = None,
        quiet_action_errors: bool = False,
        delay_step: int = 0,
    ):
        super().__init__(env, delay_step)
        self._default_pos_cfg = _ensure_dof_pattern(default_pos)
        self._offset_cfg = _ensure_dof_pattern(offset)
        self._scale_cfg = _ensure_dof_pattern(scale)
        self._clip_cfg = _ensure_dof_pattern(clip)
        self._gain_cfg = {
            "kp": _ensure_dof_pattern(pd_kp), "kv": _ensure_dof_pattern(pd_kv),
            "damping": _ensure_dof_pattern(damping), "stiffness": _ensure_dof_pattern(stiffness),
            "frictionloss": _ensure_dof_pattern(frictionloss),
        }
        self._max_force_cfg = _ensure_dof_pattern(max_force)
        self._quiet_action_errors = quiet_action_errors
        self._enabled_dof = None
        self._dofs_idx_list = None
        self._noise_scale = noise_scale
        self._reset_plan = None  # (robot, [(gain key, draw tag, engine setter)], default pose row), built on first reset
        self._fast_reset = None  # the same with everything resolved, when there is no reset noise
        self._use_default_offset = use_default_offset
        self._default_dofs_pos: torch.Tensor = None
        if use_default_offset and offset != 0.0:
            raise ValueError("Cannot set both use_default_offset and offset")
        if isinstance(joint_names, str):
            self._joint_name_cfg = [joint_names]
        elif isinstance(joint_names, list):
            self._joint_name_cfg = joint_names
        else:
            raise TypeError(f"Invalid joint_names type: {type(joint_names)}")

    # -- properties ---------------------------------------------------------------------------
    @property
    def num_actions(self) -> int:
        assert self._enabled_dof is not None, "PositionActionManager not built yet"
        return len(self._enabled_dof)

    @property
    def dofs_idx(self) -> list[int]:
        idx = self._dofs_idx_list
        if idx is None:  # (fixed once build() has resolved the joint patterns)
            idx = self._dofs_idx_list = list(self._enabled_dof.values())
        return idx

    @property
    def default_dofs_pos(self) -> torch.Tensor:
        return self._default_dofs_pos

    # -- engine getters -------------------------------------------------------------------------
    def get_dofs_position(self, noise: float = 0.0):
        if noise == 0.0:
            return self.env._trace_or("dof_pos", lambda: self.env.robot.get_dofs_position(self.dofs_idx))
        return self._add_random_noise("obs_dof_pos", self.env.robot.get_dofs_position(self.dofs_idx), noise)

    def get_dofs_velocity(self, noise: float = 0.0, clip: tuple[float, float] = None):
        if noise == 0.0 and clip is None:
            return self.env._trace_or("dof_vel", lambda: self.env.robot.get_dofs_velocity(self.dofs_idx))
        vel = self.env.robot.get_dofs_velocity(self.dofs_idx)
        if noise > 0.0:
            vel = self._add_random_noise("obs_dof_vel", vel, noise)
        if clip is not None:
            vel = vel.clamp(**clip)
        return vel

    def get_dofs_force(self, noise: float = 0.0, clip_to_max_force: bool = False):
        if noise == 0.0 and not clip_to_max_force:
            return self.env._trace_or("dof_force", lambda: self.env.robot.get_dofs_force(self.dofs_idx))
        force = self.env.robot.get_dofs_force(self.dofs_idx)
        if noise > 0.0:
            force = self._add_random_noise("obs_dof_force", force, noise)
        if clip_to_max_force and self._force_range is not None:
            force = force.clamp(self._force_range[0], self._force_range[1])
        return force

    # -- build ----------------------------------------------------------------------------------
    def build(self):
        """Resolve joint-name patterns to per-DOF parameter vectors."""
        robot = self.env.robot
        self._enabled_dof = {}
        self._dofs_idx_list = None
        for joint in robot.joints:
            if joint.type != gs.JOINT_TYPE.REVOLUTE:
                continue
            if any(re.match(f"^{p}$", joint.name) for p in self._joint_name_cfg):
                self._enabled_dof[joint.name] = joint.dof_start
        n = self.num_actions
        N = self.env.num_envs

        if self._default_pos_cfg is not None:
            default = self._dof_tensor(self._default_pos_cfg)
        else:
            default = torch.zeros(n, device=gs.device)
        self._default_dofs_pos = default.unsqueeze(0).expand(N, -1)

        lower, upper = robot.get_dofs_limit(self.dofs_idx)
        self._clip_values = torch.stack([lower, upper], dim=1).to(gs.device, gs.tc_float)
        self._scale_values = self._dof_tensor(self._scale_cfg) if self._scale_cfg is not None else None
        if self._clip_cfg is not None:
            self._dof_tensor(self._clip_cfg, output=self._clip_values)
        self._gain_values = {
            key: self._dof_tensor(cfg) for key, cfg in self._gain_cfg.items() if cfg is not None
        }
        if self._use_default_offset:
            self._offset_values = self._default_dofs_pos
        else:
            self._offset_values = self._dof_tensor(self._offset_cfg if self._offset_cfg is not None else {".*": 0.0})

        self._force_range = None
        if self._max_force_cfg is not None:
            values = self._dof_values(self._max_force_cfg)
            pairs = [v if isinstance(values[0], (list, tuple)) else (-v, v) for v in values]
            self._force_range = (
                torch.tensor([p[0] for p in pairs], device=gs.device),
                torch.tensor([p[1] for p in pairs], device=gs.device),
            )
        self._actions = torch.zeros((N, n), device=gs.device, dtype=gs.tc_float)
        self._has_stepped = False
        self._reset_plan = None
        self._fast_reset = None

    def kernel_params(self) -> dict[str, torch.Tensor]:
        """Per-DOF fp32 vectors for the action kernel: scale, offset, clip bounds, default pose."""
        offset = self._offset_values
        if offset.dim() == 2:
            offset = offset[0]
        scale = self._scale_values if self._scale_values is not None else torch.ones_like(offset)
        return {
            "scale": scale, "offset": offset,
            "clip_lo": self._clip_values[:, 0], "clip_hi": self._clip_values[:, 1],
            "default": self._default_dofs_pos[0],
        }

    # -- step / reset -----------------------------------------------------------------------------
    def step(self, actions: torch.Tensor) -> torch.Tensor:
        raise NotImplementedError(
            "action managers are stepped by ManagedEnvironment's fused step (gfb_action_step); "
            "calling step() on the manager alone is not supported"
        )

    def reset(self, envs_idx=None):
        """Engine writes for reset envs: gains (+noise) and default joint positions (+noise)."""
        if not self.enabled:
            return
        if envs_idx is None:
            envs_idx = torch.arange(self.env.num_envs, device=gs.device)
        robot = self.env.robot
        ns = self._noise_scale
        dofs_idx = self.dofs_idx
        fast = self._fast_reset
        if ns == 0.0 and fast is not None and fast[0] is robot:
            # no domain randomisation: the values handed to the engine never change -- bound setters and
            # constant rows prepared once (this runs between the step report and the re-observation
            # launch, i.e. while the GPU waits for the host)
            for setter, values in fast[1]:
                setter(values, dofs_idx, envs_idx)
            if fast[2] is not None:
                robot.set_dofs_force_range(fast[2][0], fast[2][1], dofs_idx, envs_idx)
            fast[3](position=fast[4][: envs_idx.shape[0]], dofs_idx_local=dofs_idx, envs_idx=envs_idx)
            return
        plan = self._reset_plan
        if plan is None or plan[0] is not robot:
            names = {"kp": "set_dofs_kp", "kv": "set_dofs_kv", "damping": "set_dofs_damping",
                     "stiffness": "set_dofs_stiffness", "frictionloss": "set_dofs_frictionloss"}
            tags = {"kp": "pd_kp", "kv": "pd_kv"}
            plan = self._reset_plan = (
                robot,
                [(key, f"action_dr:{tags.get(key, key)}", getattr(robot, attr))
                 for key, attr in names.items() if key in self._gain_values],
                self._default_dofs_pos[0].unsqueeze(0),
            )
        for key, tag, setter in plan[1]:
            setter(self._add_random_noise(tag, self._gain_values[key], ns), dofs_idx, envs_idx)
        if self._force_range is not None:
            lower = self._add_random_noise("action_dr:force_lower", self._force_range[0], ns)
            upper = self._add_random_noise("action_dr:force_upper", self._force_range[1], ns)
            robot.set_dofs_force_range(lower, upper, dofs_idx, envs_idx)
        # every row of the default pose is the same vector: a stride-0 view instead of an index gather
        n = envs_idx.numel() if torch.is_tensor(envs_idx) else len(envs_idx)
        fused = getattr(self.env, "_fused", None)
        if ns != 0.0 and fused is not None and not fused.dry_run and torch.is_tensor(envs_idx):
            # default pose + per-env noise for the reset envs: one launch of the reset-rows kernel
            position = fused.reset_rows("noise", "action_dr:position", envs_idx, n, len(dofs_idx), ns,
                                        base=self._default_dofs_pos[0].contiguous())
        else:
            position = self._add_random_noise("action_dr:position", plan[2].expand(n, -1), ns)
        robot.set_dofs_position(position=position, dofs_idx_local=dofs_idx, envs_idx=envs_idx)
        if ns == 0.0 and torch.is_tensor(envs_idx):
            self._fast_reset = (
                robot, [(setter, self._gain_values[key]) for key, _, setter in plan[1]], self._force_range,
                robot.set_dofs_position, self._default_dofs_pos[0].unsqueeze(0).repeat(self.env.num_envs, 1),
            )

    # -- helpers ----------------------------------------------------------------------------------
    def _dof_values(self, values: dict, default_value=0.0, output=None):
        """First matching pattern wins per DOF; a pattern that matches nothing is an error."""
        names = list(self._enabled_dof.keys())
        assigned = [False] * len(names)
        if output is None:
            output = [default_value] * len(names)
        for pattern, value in values.items():
            hit = False
            for i, name in enumerate(names):
                if assigned[i] or not re.match(f"^{pattern}$", name):
                    continue
                if isinstance(output, torch.Tensor) and not isinstance(value, torch.Tensor):
                    value = torch.tensor(value, device=gs.device)
                assigned[i], output[i], hit = True, value, True
            if not hit:
                raise RuntimeError(f"Joint DOF '{pattern}' not found.")
        return output

    def _dof_tensor(self, values: dict, default_value=0.0, output=None) -> torch.Tensor:
        out = self._dof_values(values, default_value, output)
        if isinstance(out, torch.Tensor):
            return out
        return torch.tensor(out, device=gs.device, dtype=gs.tc_float)

    def _add_random_noise(self, tag: str, values: torch.Tensor, noise_scale: float = 0.0) -> torch.Tensor:
        if noise_scale == 0.0:
            return values
        fused = getattr(self.env, "_fused", None)
        if fused is not None and not fused.dry_run and values.dim() == 1 and values.is_contiguous():
            # one row (the gains: a draw per DOF shared by all reset envs): the reset-rows kernel
            return fused.reset_rows("noise", tag, None, 1, values.shape[0], noise_scale, base=values).reshape(-1)
        return values + self.env.rng.uniform(tag, values, -1.0, 1.0) * noise_scale


class PositionWithinLimitsActionManager(PositionActionManager):
    """Actions in [-1, 1] mapped onto each joint's position limits (position_within_limits.py)."""

    kernel_mode = 2

    def __init__(
        self, env, joint_names=".*", default_pos={".*": 0.0}, pd_kp=None, pd_kv=None, max_force=None,
        damping=None, stiffness=None, frictionloss=None, noise_scale: float = 0.0, action_handler=None,
        quiet_action_errors: bool = False, delay_step: int = 0,
    ):
        super().__init__(
            env, joint_names=joint_names, default_pos=default_pos, pd_kp=pd_kp, pd_kv=pd_kv,
            max_force=max_force, damping=damping, stiffness=stiffness, frictionloss=frictionloss,
            noise_scale=noise_scale, action_handler=action_handler,
            quiet_action_errors=quiet_action_errors, delay_step=delay_step,
        )

    def build(self):
        super().build()
        lower, upper = self.env.robot.get_dofs_limit(self.dofs_idx)
        lower = lower.to(gs.device, gs.tc_float)
        upper = upper.to(gs.device, gs.tc_float)
        self._offset = (upper + lower) * 0.5
        self._scale = (upper - lower) * 0.5

    def kernel_params(self):
        p = super().kernel_params()
        p["scale"], p["offset"] = self._scale, self._offset
        return p
