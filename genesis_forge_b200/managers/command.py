"""
Command managers: per-env command vectors resampled uniformly every `resample_time_sec` and on
reset.  API of genesis_forge/managers/command/{command_manager,velocity_command}.py.

Resampling (command_manager.py:152-170, 290-303) happens inside the fused post-physics kernel; the
ranges are re-read from this object every step (they are curriculum-mutable, :99-119).  External
controllers / gamepads (command_manager.py:176-288) bypass resampling exactly as in the reference:
the `command` property then returns the controller's tensor.  The debug arrows of
VelocityCommandManager are visualisation and not part of this package.
"""
from __future__ import annotations

from typing import Callable, Tuple

import torch

from .._gs import gs
from .base import BaseManager

CommandRangeValue = Tuple[float, float]
CommandRange = CommandRangeValue | dict[str, CommandRangeValue]


class CommandManager(BaseManager):
    def __init__(self, env, range: CommandRange, resample_time_sec: float = 5.0):
        super().__init__(env, type="command")
        self._range = range
        self.resample_time_sec = resample_time_sec
        self._external_controller = None
        self._gamepad_cfg = None
        self._gamepad_axis_command_buffer = None
        num_ranges = len(range) if isinstance(range, dict) else 1
        self._command = torch.zeros(env.num_envs, num_ranges, device=gs.device)
        self._range_idx = {key: i for i, key in enumerate(range.keys())} if isinstance(range, dict) else {}

    # -- properties ---------------------------------------------------------------------------
    @property
    def command(self) -> torch.Tensor:
        if self._external_controller is not None:
            return self._external_controller(self.env.step_count)
        return self._command

    @property
    def range(self) -> CommandRange:
        return self._range

    @range.setter
    def range(self, range: CommandRange):
        num = len(range) if isinstance(range, dict) else 1
        if num != self._command.shape[1]:
            raise ValueError(
                f"Cannot change the shape of the CommandManager range. Expected size: {self._command.shape[1]}, got {num}"
            )
        if type(range) != type(self._range):
            raise ValueError(
                f"Cannot change the base type of the CommandManager range. Expected type: {type(self._range)}, got {type(range)}"
            )
        if isinstance(range, dict) and set(range.keys()) != set(self._range.keys()):
            raise ValueError(
                f"Cannot change the dict keys of the CommandManager range. Expected keys: {set(self._range.keys())}, got {set(range.keys())}"
            )
        self._range = range

    @property
    def resample_time_sec(self) -> float:
        return self._resample_time_sec

    @resample_time_sec.setter
    def resample_time_sec(self, resample_time_sec: float):
        self._resample_time_sec = resample_time_sec
        self._resample_steps = int(resample_time_sec / self.env.dt)

    def ranges_list(self) -> list[CommandRangeValue]:
        return list(self._range.values()) if isinstance(self._range, dict) else [self._range]

    # -- operations ---------------------------------------------------------------------------
    def get_command(self, key: str) -> torch.Tensor:
        if not isinstance(self._range, dict):
            raise ValueError("The range is not a dict")
        return self._command[:, self._range_idx[key]]

    def get_command_idx(self, key: str) -> int:
        if not isinstance(self._range, dict):
            raise ValueError("The range is not a dict")
        return self._range_idx[key]

    # -- lifecycle -------------------------------------------------------------------------------
    # The stock managers (this class and VelocityCommandManager, neither method overridden) never get
    # here: their interval / reset resampling is part of the fused post-physics kernel.  A SUBCLASS
    # that overrides step(), reset() or resample_command() -- the way examples/gait_trainer's
    # GaitCommandManager does -- is stepped and reset on the host by ManagedEnvironment, and its
    # super().step() / super().reset(env_ids) calls land here: the host-side equivalents of
    # command_manager.py:152-170, resampling through the (possibly overridden) resample_command.
    def step(self):
        if not self.enabled or self._external_controller is not None:
            return
        # the in-library reset runs on the host side of this call in split execution, but the
        # interval test must see the episode lengths of BEFORE the reset, as in the reference
        # (managed_env.py:318-323): ManagedEnvironment steps these managers ahead of the reset phase
        due = (self.env.episode_length % self._resample_steps == 0).nonzero(as_tuple=False).reshape((-1,))
        self.resample_command(due)

    def reset(self, env_ids=None):
        if not self.enabled:
            return
        if env_ids is None:
            env_ids = torch.arange(self.env.num_envs, device=gs.device)
        self.resample_command(env_ids)

    def observation(self, env) -> torch.Tensor:
        """Observation term: the current command (command_manager.py:172-174)."""
        return self.env._trace_or(("command", self), lambda: self.command)

    def use_external_controller(self, controller: Callable[[int], torch.Tensor]):
        self._external_controller = controller
        self._notify_fused()

    def _notify_fused(self):
        fused = getattr(self.env, "_fused", None)
        if fused is not None:
            fused._has_command_override = True  # the fused step re-binds this manager's command source

    def use_gamepad(self, gamepad, range_axis: int | dict[str, int]):
        self._external_controller = self._gamepad_axis_command
        axis_map = []
        if isinstance(range_axis, int):
            axis_map.append(range_axis)
        elif isinstance(range_axis, dict):
            axis_map = [range_axis[key] for key in self._range.keys()]
        self._gamepad_cfg = {"gamepad": gamepad, "axis_map": axis_map}
        self._gamepad_axis_command_buffer = torch.zeros_like(self._command, device=gs.device)
        self._notify_fused()

    def resample_command(self, env_ids):
        """Draw a new command for `env_ids` now (host path; the per-step resample is in-kernel)."""
        ranges = self.ranges_list()
        for i in range(self._command.shape[1]):
            like = torch.empty(len(env_ids), device=gs.device)
            self._command[env_ids, i] = self.env.rng.uniform(f"cmd_host:{i}", like, *ranges[i])

    def _gamepad_axis_command(self, step_count: int) -> torch.Tensor:
        if self._gamepad_cfg is None:
            return self._gamepad_axis_command_buffer
        pad, axis_map = self._gamepad_cfg["gamepad"], self._gamepad_cfg["axis_map"]
        cmd = self._gamepad_axis_command_buffer
        for i, axis in enumerate(axis_map):
            if i < len(self.ranges_list()):
                lo, hi = self.ranges_list()[i]
                cmd[:, i] = (pad.state.axis(axis) + 1.0) * (hi - lo) / 2 + lo
        return cmd


class VelocityCommandManager(CommandManager):
    """
    (lin_vel_x, lin_vel_y, ang_vel_z) command in the robot frame.

    `standing_probability` is accepted and stored but has NO effect, as in the reference: its
    `_resample_command` override (velocity_command.py:194-207) is never called because the base
    class calls `resample_command` (command_manager.py:162,170).
    """

    def __init__(
        self, env, range, resample_time_sec: float = 5.0, standing_probability: float = 0.0,
        debug_visualizer: bool = False, debug_visualizer_cfg: dict | None = None,
    ):
        super().__init__(env, range=range, resample_time_sec=resample_time_sec)
        self.standing_probability = standing_probability
        self.debug_visualizer = debug_visualizer
        self.visualizer_cfg = dict(debug_visualizer_cfg or {})
        self._is_standing_env = torch.zeros(env.num_envs, dtype=torch.bool, device=gs.device)

    def use_gamepad(self, gamepad, lin_vel_y_axis: int = 0, lin_vel_x_axis: int = 1, ang_vel_z_axis: int = 2):
        super().use_gamepad(
            gamepad,
            range_axis={"lin_vel_x": lin_vel_x_axis, "lin_vel_y": lin_vel_y_axis, "ang_vel_z": ang_vel_z_axis},
        )
