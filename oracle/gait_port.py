"""
TEST INFRASTRUCTURE (oracle).  torch-CPU restatement of the gait_trainer example's command manager
(/root/reference/examples/gait_trainer/gait_command_manager.py), the user-level manager of
BASELINE config 3b.  Same role and rules as oracle/manager_port.py: every eager op sequence follows
the cited lines (same op order, same python-scalar operands, same draw order from torch's global
CPU generator), so that the bits are the example's bits; tests/test_oracle_vs_reference.py pins it
against the UNMODIFIED example class imported from /root/reference in the build container.

Never imported by the product package.
"""
from __future__ import annotations

import torch

# gait_command_manager.py:26-63 (walk and canter are commented out in the example)
GAIT_OFFSETS = {
    "trot": {"FL": 0.0, "FR": 0.5, "RL": 0.5, "RR": 0.0},
    "pace": {"FL": 0.5, "FR": 0.0, "RL": 0.5, "RR": 0.0},
    "bound": {"FL": 0.0, "FR": 0.0, "RL": 0.5, "RR": 0.5},
    "pronk": {"FL": 0.0, "FR": 0.0, "RL": 0.0, "RR": 0.0},
}
GAIT_PERIOD_RANGE = [0.3, 0.6]          # :17
FOOT_CLEARANCE_RANGE = [0.04, 0.12]     # :18


class GaitPort:
    def __init__(self, env, cfg: dict):
        """gait_command_manager.py:86-132; `env` is the PortEnv."""
        self.env = env
        N = env.num_envs
        self.resample_steps = int(cfg["resample_time_sec"] / env.dt)  # command_manager.py:127-130
        self.foot_names = cfg["foot_names"]
        self.foot_links = []
        self._command = torch.zeros(N, 0)  # CommandManager built with range={} (command_manager.py:75-78)
        self.num_gaits = 1
        self.gait_period_range = [(GAIT_PERIOD_RANGE[0] + GAIT_PERIOD_RANGE[1]) / 2] * 2
        self.foot_clearance_range = [FOOT_CLEARANCE_RANGE[0]] * 2
        self.all_gaits_learned = False
        self.foot_offset = torch.zeros((N, 4))
        self.gait_period = torch.zeros((N, 1))
        self.foot_height = torch.zeros((N, 1))
        self.gait_time = torch.zeros(N, 1, dtype=torch.float)
        self.gait_phase = torch.zeros(N, 1, dtype=torch.float)
        self.clock_input = torch.zeros(N, 8, dtype=torch.float)
        self.gait_selected = torch.zeros(N, dtype=torch.long)
        for what, times in (cfg.get("curriculum") or {}).items():
            for _ in range(times):
                getattr(self, f"increment_{what}")()

    def build(self):
        """:213-222"""
        for i, key in enumerate(("FL", "FR", "RL", "RR")):
            self.foot_links.insert(i, self.env.robot.get_link(self.foot_names[key]))

    # -- curriculum (:150-190) ---------------------------------------------------------------------
    def increment_num_gaits(self):
        if self.all_gaits_learned:
            return
        if self.num_gaits == len(GAIT_OFFSETS):
            self.all_gaits_learned = True
        else:
            self.num_gaits = min(self.num_gaits + 1, len(GAIT_OFFSETS))

    def increment_gait_period_range(self):
        self.gait_period_range[0] = max(self.gait_period_range[0] - 0.05, GAIT_PERIOD_RANGE[0])
        self.gait_period_range[1] = min(self.gait_period_range[1] + 0.05, GAIT_PERIOD_RANGE[1])

    def increment_foot_clearance_range(self):
        self.foot_clearance_range[0] = max(self.foot_clearance_range[0] - 0.01, FOOT_CLEARANCE_RANGE[0])
        self.foot_clearance_range[1] = min(self.foot_clearance_range[1] + 0.01, FOOT_CLEARANCE_RANGE[1])

    # -- command / observation (:134-148, :243-255) -------------------------------------------------
    @property
    def command(self) -> torch.Tensor:
        return torch.cat([self.foot_offset, self.foot_height, self.gait_period], dim=-1)

    def observation(self) -> torch.Tensor:
        return torch.cat([self.command, self.clock_input], dim=-1)

    # -- resampling (:192-221, :356-399) --------------------------------------------------------------
    def resample_command(self, env_ids: torch.Tensor):
        gait_names = list(GAIT_OFFSETS.keys())[: self.num_gaits]
        if self.num_gaits == 1:
            self._set_gait(gait_names[0], env_ids)
            self.gait_selected[env_ids] = 0
        else:
            gait_indices = self._generate_random_gait_indices(len(env_ids))
            for gait_idx in range(self.num_gaits):
                mask = gait_indices == gait_idx
                if mask.any():
                    selected_envs = env_ids[mask]
                    self._set_gait(gait_names[gait_idx], selected_envs)
                    self.gait_selected[selected_envs] = gait_idx

    def _set_gait(self, gait_name: str, env_ids: torch.Tensor):
        log = self.env.rng_log
        gait_offsets = GAIT_OFFSETS[gait_name]
        self.foot_offset[env_ids, 0] = gait_offsets["FL"]
        self.foot_offset[env_ids, 1] = gait_offsets["FR"]
        self.foot_offset[env_ids, 2] = gait_offsets["RL"]
        self.foot_offset[env_ids, 3] = gait_offsets["RR"]
        if gait_name in ["pronk", "bound"]:
            self.foot_height[env_ids, 0] = self.foot_clearance_range[0]
        else:
            draw = torch.empty(len(env_ids)).uniform_(*self.foot_clearance_range)
            log.append(("gait_height", draw.clone()))
            self.foot_height[env_ids, 0] = draw
        draw = torch.empty(len(env_ids)).uniform_(*self.gait_period_range)
        log.append(("gait_period", draw.clone()))
        self.gait_period[env_ids, 0] = draw

    def _generate_random_gait_indices(self, num: int) -> torch.Tensor:
        if not self.all_gaits_learned:
            weights = torch.arange(self.num_gaits).exp()
        else:
            weights = torch.ones(self.num_gaits)
        weights /= weights.sum()
        weights = weights[: self.num_gaits].expand(num, -1)
        picks = torch.multinomial(weights, 1).squeeze(-1)
        self.env.rng_log.append(("gait_pick", picks.clone()))
        return picks

    # -- lifecycle (:224-241, command_manager.py:152-170) -----------------------------------------------
    def step(self):
        env = self.env
        idx = (env.episode_length % self.resample_steps == 0).nonzero(as_tuple=False).reshape((-1,))
        self.resample_command(idx)
        # _log_metrics (:446-455)
        logging = env.extras["episode"]
        logging["Metrics / num_gaits"] = self.num_gaits
        for i, gait_name in enumerate(GAIT_OFFSETS.keys()):
            logging[f"Metrics / gait_{gait_name}_envs"] = (self.gait_selected == i).sum()
        self.gait_time = (self.gait_time + env.dt) % self.gait_period
        self.gait_phase = self.gait_time / self.gait_period
        for i in range(4):
            foot_phase = (self.gait_phase + self.foot_offset[:, i].unsqueeze(1)) % 1.0
            self.clock_input[:, i] = torch.sin(2 * torch.pi * foot_phase).squeeze(-1)
            self.clock_input[:, i + 4] = torch.cos(2 * torch.pi * foot_phase).squeeze(-1)

    def reset(self, env_ids):
        if env_ids is None:
            env_ids = torch.arange(self.env.num_envs)
        self.resample_command(env_ids)
        self.clock_input[env_ids, :] = 0.0
        self.gait_time[env_ids] = 0.0
        self.gait_phase[env_ids] = 0.0

    # -- reward terms (:257-345) ------------------------------------------------------------------------
    def foot_height_reward(self, sensitivity: float = 0.1) -> torch.Tensor:
        robot = self.env.robot
        link_idx = [f.idx_local for f in self.foot_links]
        foot_vel = robot.get_links_vel(links_idx_local=link_idx)
        foot_pos = robot.get_links_pos(links_idx_local=link_idx)
        foot_vel_xy_norm = torch.norm(foot_vel[:, :, :2], dim=-1)
        clearance_error = torch.sum(foot_vel_xy_norm * torch.square(foot_pos[:, :, 2] - self.foot_height), dim=-1)
        return torch.exp(-clearance_error / sensitivity)

    def gait_phase_reward(self, contact: dict) -> torch.Tensor:
        fl = self._foot_phase_reward(0, contact)
        fr = self._foot_phase_reward(1, contact)
        rl = self._foot_phase_reward(2, contact)
        rr = self._foot_phase_reward(3, contact)
        quad_reward = fl.flatten() + fr.flatten() + rl.flatten() + rr.flatten()
        return torch.exp(quad_reward)

    def _foot_phase_reward(self, foot_idx: int, contact: dict) -> torch.Tensor:
        N = self.env.num_envs
        link = self.foot_links[foot_idx]
        force_weight = torch.zeros(N, 1, dtype=torch.float)
        vel_weight = torch.zeros(N, 1, dtype=torch.float)
        # contact_manager.py:258-269
        idx = torch.nonzero(contact["link_ids"] == link.idx)[0]
        force = torch.norm(contact["contacts"][:, idx, :], dim=-1).view(-1, 1)
        velocity = torch.norm(link.get_vel(), dim=-1).view(-1, 1)
        phi = (self.gait_phase + self.foot_offset[:, foot_idx].unsqueeze(1)) % 1.0
        phi *= 2 * torch.pi
        swing_indices = (phi >= 0.0) & (phi < torch.pi)
        swing_indices = swing_indices.nonzero().flatten()
        stance_indices = (phi >= torch.pi) & (phi < 2 * torch.pi)
        stance_indices = stance_indices.nonzero().flatten()
        force_weight[swing_indices, :] = -1
        vel_weight[swing_indices, :] = 0
        force_weight[stance_indices, :] = 0
        vel_weight[stance_indices, :] = -1
        return vel_weight * velocity + force_weight * force

    def snapshot(self, name: str) -> dict:
        return {
            f"gait/{name}/foot_offset": self.foot_offset, f"gait/{name}/gait_period": self.gait_period,
            f"gait/{name}/foot_height": self.foot_height, f"gait/{name}/gait_time": self.gait_time,
            f"gait/{name}/gait_phase": self.gait_phase, f"gait/{name}/clock_input": self.clock_input,
            f"gait/{name}/gait_selected": self.gait_selected,
        }
