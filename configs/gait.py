"""
User-level command manager of the gait_trainer workload (BASELINE config 3, second half).

In the reference this manager is NOT library code: it lives in the example
(examples/gait_trainer/gait_command_manager.py) and extends `CommandManager` the way a user would --
overriding `command`, `step`, `reset`, `resample_command`, `observation` and bringing two reward
terms of its own (`gait_phase_reward` :253-266, `foot_height_reward` :234-251).  This module is the
same kind of user code written against the manager API that the reference and the drop-in share, so
`make_gait_command_manager(ns.managers.CommandManager)` works on either.  It is a workload
definition (like configs/specs.py); the arithmetic it contains is user code that the drop-in runs as
host callbacks between kernel phases (split execution), or -- when the fused gait terms are enabled
-- recognises and lowers to kernel opcodes.

Periodic-reward gait model (Siekmann et al. 2020): every env carries a gait period, a swing height
and one phase offset per foot; a clock `gait_time` advances by dt modulo the period; a foot whose
phase (clock phase + offset, mod 1) is in the first half of the cycle is swinging (contact force
penalised), otherwise in stance (foot speed penalised).

Random draws go through `env.rng` when the environment has one (the drop-in: lets the parity harness
replay the oracle's draws), else through torch's global generator like the reference example does.
"""
from __future__ import annotations

import math

import torch

# phase offset per foot (FL, FR, RL, RR); insertion order is the curriculum order (:32-63)
GAITS = {
    "trot": (0.0, 0.5, 0.5, 0.0),
    "pace": (0.5, 0.0, 0.5, 0.0),
    "bound": (0.0, 0.0, 0.5, 0.5),
    "pronk": (0.0, 0.0, 0.0, 0.0),
}
LOW_CLEARANCE_GAITS = ("pronk", "bound")
PERIOD_LIMITS = (0.3, 0.6)
CLEARANCE_LIMITS = (0.04, 0.12)
FEET = ("FL", "FR", "RL", "RR")


def make_gait_command_manager(command_manager_cls):
    """The gait manager as a subclass of the given namespace's CommandManager."""

    class GaitCommandManager(command_manager_cls):
        def __init__(self, env, foot_names: dict, resample_time_sec: float = 5.0, robot_entity_attr: str = "robot"):
            super().__init__(env, range={}, resample_time_sec=resample_time_sec)
            self._robot_entity_attr = robot_entity_attr
            self._foot_names = foot_names
            self.foot_links = []
            dev, n = self._command.device, env.num_envs
            # curriculum state: one gait, the middle period, the lowest clearance to begin with
            self._num_gaits = 1
            self._gait_period_range = [sum(PERIOD_LIMITS) / 2] * 2
            self._foot_clearance_range = [CLEARANCE_LIMITS[0]] * 2
            self._all_gaits_learned = False
            self.foot_offset = torch.zeros((n, 4), device=dev)
            self.gait_period = torch.zeros((n, 1), device=dev)
            self.foot_height = torch.zeros((n, 1), device=dev)
            self.gait_time = torch.zeros((n, 1), device=dev)
            self.gait_phase = torch.zeros((n, 1), device=dev)
            self.clock_input = torch.zeros((n, 8), device=dev)
            self._gait_selected = torch.zeros(n, dtype=torch.long, device=dev)

        # -- the command: [4 foot offsets, swing height, period] ---------------------------------
        @property
        def command(self) -> torch.Tensor:
            return torch.cat([self.foot_offset, self.foot_height, self.gait_period], dim=-1)

        def observation(self, env) -> torch.Tensor:
            return torch.cat([self.command, self.clock_input], dim=-1)

        # -- curriculum (:150-190) ---------------------------------------------------------------
        def increment_num_gaits(self):
            if self._all_gaits_learned:
                return
            if self._num_gaits == len(GAITS):
                self._all_gaits_learned = True
            else:
                self._num_gaits = min(self._num_gaits + 1, len(GAITS))

        def increment_gait_period_range(self):
            lo, hi = self._gait_period_range
            self._gait_period_range = [max(lo - 0.05, PERIOD_LIMITS[0]), min(hi + 0.05, PERIOD_LIMITS[1])]

        def increment_foot_clearance_range(self):
            lo, hi = self._foot_clearance_range
            self._foot_clearance_range = [max(lo - 0.01, CLEARANCE_LIMITS[0]), min(hi + 0.01, CLEARANCE_LIMITS[1])]

        # -- draws -------------------------------------------------------------------------------
        def _uniform(self, tag: str, n: int, lo: float, hi: float) -> torch.Tensor:
            like = torch.empty(n, device=self.foot_offset.device)
            rng = getattr(self.env, "rng", None)
            return rng.uniform(tag, like, lo, hi) if rng is not None else like.uniform_(lo, hi)

        def _pick_gaits(self, n: int) -> torch.Tensor:
            """Recently unlocked gaits are exponentially more likely until all are learned (:381-399)."""
            dev = self.foot_offset.device
            if self._all_gaits_learned:
                weights = torch.ones(self._num_gaits, device=dev)
            else:
                weights = torch.arange(self._num_gaits, device=dev).exp()
            weights = (weights / weights.sum()).expand(n, -1)
            rng = getattr(self.env, "rng", None)
            if rng is not None and hasattr(rng, "multinomial"):
                return rng.multinomial("gait_pick", weights)
            return torch.multinomial(weights, 1).squeeze(-1)

        # -- lifecycle ---------------------------------------------------------------------------
        def build(self):
            super().build()
            robot = getattr(self.env, self._robot_entity_attr)
            self.foot_links = [robot.get_link(self._foot_names[key]) for key in FEET]

        def resample_command(self, env_ids):
            if isinstance(env_ids, list):
                env_ids = torch.tensor(env_ids, device=self.foot_offset.device, dtype=torch.long)
            names = list(GAITS)[: self._num_gaits]
            if self._num_gaits == 1:
                self._assign(names[0], env_ids)
                self._gait_selected[env_ids] = 0
                return
            picks = self._pick_gaits(len(env_ids))
            for g, name in enumerate(names):
                chosen = picks == g
                if chosen.any():
                    self._assign(name, env_ids[chosen])
                    self._gait_selected[env_ids[chosen]] = g

        def _assign(self, gait: str, env_ids: torch.Tensor):
            for foot, offset in enumerate(GAITS[gait]):
                self.foot_offset[env_ids, foot] = offset
            if gait in LOW_CLEARANCE_GAITS:
                self.foot_height[env_ids, 0] = self._foot_clearance_range[0]
            else:
                self.foot_height[env_ids, 0] = self._uniform("gait_height", len(env_ids), *self._foot_clearance_range)
            self.gait_period[env_ids, 0] = self._uniform("gait_period", len(env_ids), *self._gait_period_range)

        def step(self):
            super().step()  # resamples on the interval through resample_command above
            log = self.env.extras[self.env.extras_logging_key]
            log["Metrics / num_gaits"] = self._num_gaits
            for g, name in enumerate(GAITS):
                log[f"Metrics / gait_{name}_envs"] = (self._gait_selected == g).sum()
            # the clock: time modulo period, and per foot sin / cos of its phase
            self.gait_time = (self.gait_time + self.env.dt) % self.gait_period
            self.gait_phase = self.gait_time / self.gait_period
            angle = 2 * torch.pi * ((self.gait_phase + self.foot_offset) % 1.0)
            self.clock_input[:, :4] = torch.sin(angle)
            self.clock_input[:, 4:] = torch.cos(angle)

        def reset(self, env_ids=None):
            if env_ids is None:
                env_ids = torch.arange(self.env.num_envs, device=self.foot_offset.device)
            super().reset(env_ids)
            self.clock_input[env_ids, :] = 0.0
            self.gait_time[env_ids] = 0.0
            self.gait_phase[env_ids] = 0.0

        # -- reward terms ------------------------------------------------------------------------
        def foot_height_reward(self, env, sensitivity: float = 0.1) -> torch.Tensor:
            """exp(-sum_feet |v_xy| (z - swing height)^2 / sensitivity): feet reach the height while moving."""
            local = [link.idx_local for link in self.foot_links]
            vel = env.robot.get_links_vel(links_idx_local=local)
            pos = env.robot.get_links_pos(links_idx_local=local)
            speed = torch.norm(vel[:, :, :2], dim=-1)
            error = torch.sum(speed * torch.square(pos[:, :, 2] - self.foot_height), dim=-1)
            return torch.exp(-error / sensitivity)

        def gait_phase_reward(self, env, contact_manager) -> torch.Tensor:
            """exp(sum_feet -(swing ? |contact force| : |foot velocity|))."""
            total = None
            for foot, link in enumerate(self.foot_links):
                force = torch.norm(contact_manager.get_contact_forces(link.idx), dim=-1).view(-1, 1)
                speed = torch.norm(link.get_vel(), dim=-1).view(-1, 1)
                phi = (self.gait_phase + self.foot_offset[:, foot].unsqueeze(1)) % 1.0
                phi = phi * (2 * torch.pi)
                # the example selects rows with `.nonzero().flatten()` of an (N, 1) mask (:316-319): the
                # flattened (row, col) pairs also contain the column index 0, so env 0 is written whenever
                # a set is non-empty -- stance last.  Same here, or env 0 would differ from the reference.
                swing_rows = ((phi >= 0.0) & (phi < torch.pi)).nonzero().flatten()
                stance_rows = ((phi >= torch.pi) & (phi < 2 * torch.pi)).nonzero().flatten()
                force_weight = torch.zeros_like(phi)
                speed_weight = torch.zeros_like(phi)
                force_weight[swing_rows, :] = -1
                force_weight[stance_rows, :] = 0
                speed_weight[swing_rows, :] = 0
                speed_weight[stance_rows, :] = -1
                term = (speed_weight * speed + force_weight * force).flatten()
                total = term if total is None else total + term
            return torch.exp(total)

    return GaitCommandManager


def reference_gait_command_manager():
    """
    The UNMODIFIED example class (build container only): examples/gait_trainer/gait_command_manager.py
    imported from /root/reference under the oracle's stub modules.
    """
    import importlib.util
    import sys

    from oracle import shim

    shim.import_reference()
    name = "gait_command_manager"
    if name not in sys.modules:
        path = shim.REFERENCE_ROOT + "/examples/gait_trainer/gait_command_manager.py"
        spec = importlib.util.spec_from_file_location(name, path)
        module = importlib.util.module_from_spec(spec)
        sys.modules[name] = module
        spec.loader.exec_module(module)
    return sys.modules[name].GaitCommandManager


assert math.isclose(sum(PERIOD_LIMITS) / 2, 0.45)
