"""
TEST INFRASTRUCTURE (oracle) -- the oracle proper.

A torch-CPU, op-for-op restatement of everything `ManagedEnvironment.step / reset / build` does in
the reference (genesis_forge/managed_env.py:249-398), driven by a term-table spec (configs/specs.py)
instead of manager objects.  It exists because the reference itself (pure Python) cannot travel to
the GPU box; this file can.  It is pinned, bit for bit and step by step (every manager buffer, every
output, every extras entry, and the global torch RNG stream), against the UNMODIFIED reference on
all spec configs by tests/test_oracle_vs_reference.py in the build container, and against the
committed golden traces generated from the reference (tests/golden/) everywhere.

Every eager op sequence below follows the cited reference lines exactly (same op order, same
in-place/out-of-place choice, same python-scalar operands) so that the torch-CPU bits are the
reference's bits.  The third-party quaternion helpers come from oracle/geom.py (parity unpinned).

Never imported by the product package.  bench.py times it as the `cpu_baseline` ("port").
"""
from __future__ import annotations

import math
import re

import torch
import torch.nn.functional as F

from .contact_kernel import kernel_get_contact_forces
from .geom import inv_quat, transform_by_quat, xyz_to_quat


class _ContactView:
    """What user-defined terms may read of a contact manager (same attribute names as the reference)."""

    def __init__(self, m):
        self._m = m

    contacts = property(lambda self: self._m["contacts"])
    contact_positions = property(lambda self: self._m["positions"])
    last_air_time = property(lambda self: self._m.get("last_air"))
    current_air_time = property(lambda self: self._m.get("cur_air"))
    last_contact_time = property(lambda self: self._m.get("last_contact"))
    current_contact_time = property(lambda self: self._m.get("cur_contact"))


class _CommandView:
    def __init__(self, c):
        self._c = c

    _command = property(lambda self: self._c["command"])
    command = property(lambda self: self._c["command"])


class PortEnv:
    """Spec-driven restatement of ManagedEnvironment + all managers.  CPU tensors only."""

    def __init__(self, spec: dict, num_envs: int, scene, terrain, robot, record_margins: bool = False):
        self.spec = spec
        self.num_envs = N = num_envs
        self.dt = spec["dt"]
        self.scene, self.terrain, self.robot = scene, terrain, robot
        self.device = torch.device("cpu")
        self.rng_log: list[tuple[str, torch.Tensor]] = []   # (tag, draw) in generation order
        self.margins: list[tuple[str, torch.Tensor, float]] = []  # (what, value, threshold)
        self.record_margins = record_margins
        self.printed: list[str] = []
        self.body_acc_state: dict[str, dict] = {}
        # manager kinds whose `enabled` flag is off ("action", "termination", "reward"): their step()
        # returns before touching anything (position_action_manager.py:383-384,
        # termination_manager.py:159-160, reward_manager.py:172-173)
        self.disabled: set[str] = set()

        # genesis_env.py:60-93
        self.extras = {"episode": {}}
        self.actions = None
        self.last_actions = None
        self.step_count = 0
        self.episode_length = torch.zeros((N,), dtype=torch.int32)
        self.max_episode_length = None
        self.base_max_episode_length = None
        self.max_episode_random_scaling = spec.get("max_episode_random_scaling", 0.0)
        sec = spec.get("max_episode_length_sec", 10)
        if sec and sec > 0:
            self.max_episode_length = torch.zeros((N,), dtype=torch.int32)
            self.base_max_episode_length = math.ceil(sec / self.dt)  # genesis_env.py:166
            self.max_episode_length[:] = self.base_max_episode_length

        if "fixed_command" in spec:
            self.fixed_command = torch.zeros((N, 3), dtype=torch.float32)
            for i, v in enumerate(spec["fixed_command"]):
                self.fixed_command[:, i] = v

    # ------------------------------------------------------------------------------------------
    # helpers
    # ------------------------------------------------------------------------------------------
    def _uniform(self, tag: str, shape, lo: float, hi: float) -> torch.Tensor:
        t = torch.empty(shape).uniform_(lo, hi)
        self.rng_log.append((tag, t.clone()))
        return t

    def _draw_max_len(self, idx: torch.Tensor) -> torch.Tensor:
        """U(-1, 1) draws of genesis_env.py:249 (a hook: the parity harness can transplant the kernel's draws)."""
        return self._uniform("max_len", (idx.numel(),), -1.0, 1.0)

    def _margin(self, what: str, value: torch.Tensor, threshold: float):
        if self.record_margins:
            self.margins.append((what, value.detach().clone(), float(threshold)))

    def _links_idx(self, entity, names):
        """contact_manager.py:342-382"""
        ids, local = [], []
        if names is None:
            for link in entity.links:
                ids.append(link.idx)
                local.append(link.idx_local)
        else:
            for pattern in names:
                found = False
                for link in entity.links:
                    if pattern == link.name or re.match(f"^{pattern}$", link.name):
                        ids.append(link.idx)
                        local.append(link.idx_local)
                        found = True
                if not found:
                    raise RuntimeError(f"Link '{pattern}' not found")
        return torch.tensor(ids), torch.tensor(local)

    def _dof_values(self, cfg, default=0.0, output=None):
        """position_action_manager.py:470-514 (first matching pattern wins per DOF)."""
        names = list(self.enabled_dof.keys())
        is_set = [False] * len(names)
        if output is None:
            output = [default] * len(names)
        for pattern, value in cfg.items():
            found = False
            for i, name in enumerate(names):
                if not is_set[i] and re.match(f"^{pattern}$", name):
                    if isinstance(output, torch.Tensor) and not isinstance(value, torch.Tensor):
                        value = torch.tensor(value)
                    is_set[i] = True
                    output[i] = value
                    found = True
            if not found:
                raise RuntimeError(f"Joint DOF '{pattern}' not found.")
        return output

    def _dof_tensor(self, cfg, default=0.0, output=None):
        return torch.tensor(self._dof_values(cfg, default, output), dtype=torch.float32)

    @staticmethod
    def _pattern(value):
        if value is None:
            return None
        return value if isinstance(value, dict) else {".*": value}

    # ------------------------------------------------------------------------------------------
    # build  (managed_env.py:249-272; order terrain, action, contact, termination, reward,
    #         command, entity, observation)
    # ------------------------------------------------------------------------------------------
    def build(self):
        N, spec = self.num_envs, self.spec
        self.scene.build(n_envs=N)

        # --- terrain_manager.py:281-359
        self.has_terrain = "terrain" in spec
        if self.has_terrain:
            morph = self.terrain.morph
            self.t_origin = morph.pos
            sx = morph.subterrain_size[0] * morph.n_subterrains[0]
            sy = morph.subterrain_size[1] * morph.n_subterrains[1]
            self.t_size = (sx, sy)
            x_min, y_min = self.t_origin[0], self.t_origin[1]
            self.t_bounds = (x_min, x_min + sx, y_min, y_min + sy)
            hf = torch.as_tensor(self.terrain.geoms[0].metadata["height_field"], dtype=torch.float32)
            hf = hf * morph.vertical_scale
            self.t_height_field = hf.T
            self.t_norm = torch.zeros((N, 2))
            self.t_grid = torch.zeros((N, 1, 1, 2))
            self.t_heights = torch.zeros(N)
            self.t_env_pos = torch.zeros((N, 3))

        # --- position_action_manager.py:297-374
        a = spec["action"]
        self.enabled_dof = {}
        patterns = a["joint_names"] if isinstance(a["joint_names"], list) else [a["joint_names"]]
        for joint in self.robot.joints:
            if joint.type.name != "REVOLUTE":
                continue
            for pattern in patterns:
                if re.match(f"^{pattern}$", joint.name):
                    self.enabled_dof[joint.name] = joint.dof_start
                    break
        self.dofs_idx = list(self.enabled_dof.values())
        D = self.num_actions = len(self.dofs_idx)
        self.default_dofs_pos = self._dof_tensor(self._pattern(a.get("default_pos", {".*": 0.0})))
        self.default_dofs_pos = self.default_dofs_pos.unsqueeze(0).expand(N, -1)
        lower, upper = self.robot.get_dofs_limit(self.dofs_idx)
        self.clip_values = torch.stack([lower, upper], dim=1)
        self.scale_values = self._dof_tensor(self._pattern(a.get("scale", 1.0)))
        if a.get("clip") is not None:
            self._dof_tensor(self._pattern(a["clip"]), output=self.clip_values)
        self.gain_values = {}
        for key in ("pd_kp", "pd_kv", "damping", "stiffness", "frictionloss"):
            if a.get(key) is not None:
                self.gain_values[key] = self._dof_tensor(self._pattern(a[key]))
        if a.get("use_default_offset", True):
            self.offset_values = self.default_dofs_pos
        else:
            self.offset_values = self._dof_tensor(self._pattern(a.get("offset", 0.0)))
        self.force_range = None
        if a.get("max_force") is not None:
            mf = self._dof_values(self._pattern(a["max_force"]))
            lo, hi = [0.0] * D, [0.0] * D
            for i, v in enumerate(mf):
                if isinstance(mf[0], (list, tuple)):
                    lo[i], hi[i] = v[0], v[1]
                else:
                    lo[i], hi[i] = -v, v
            self.force_range = (torch.tensor(lo), torch.tensor(hi))
        self.action_noise_scale = a.get("noise_scale", 0.0)
        self.delay_step = a.get("delay_step", 0)
        self.action_delay_buffer = []
        self.within_limits = a["type"] == "within_limits"
        if self.within_limits:  # position_within_limits.py:99-111
            lo_e = lower.unsqueeze(0).expand(N, -1)
            hi_e = upper.unsqueeze(0).expand(N, -1)
            self.wl_offset = (hi_e + lo_e) * 0.5
            self.wl_scale = (hi_e - lo_e) * 0.5
        self.targets = None      # action manager's _actions
        self.raw_actions = None

        # --- contact_manager.py:271-314
        self.contact = {}
        for name, c in spec["contacts"].items():
            entity = getattr(self, c.get("entity_attr", "robot"))
            link_ids, local_ids = self._links_idx(entity, c["link_names"])
            with_ids = torch.empty(0)
            has_filter = c.get("with_entity_attr") is not None or c.get("with_links_names") is not None
            if c.get("with_entity_attr") or c.get("with_links_names"):
                w_entity = getattr(self, c.get("with_entity_attr") or "robot")
                with_ids, _ = self._links_idx(w_entity, c.get("with_links_names"))
            Lc = link_ids.shape[0]
            m = {
                "link_ids": link_ids, "local_link_ids": local_ids, "with_link_ids": with_ids,
                "has_filter": has_filter, "track": c.get("track_air_time", False),
                "threshold": c.get("air_time_contact_threshold", 1.0),
                "contacts": torch.zeros((N, Lc, 3)), "positions": torch.zeros((N, Lc, 3)),
                "counts": torch.zeros((N, Lc)),
            }
            if m["track"]:
                for k in ("last_air", "cur_air", "last_contact", "cur_contact"):
                    m[k] = torch.zeros((N, Lc))
            self.contact[name] = m
            setattr(self, name, _ContactView(m))

        # --- termination_manager.py:116-119
        self.terminated = torch.zeros(N, dtype=torch.bool)
        self.truncated = torch.zeros(N, dtype=torch.bool)

        # --- reward_manager.py:107-118
        self.reward_buf = torch.zeros((N,))
        self.episode_seconds = torch.zeros((N,))
        self.episode_mean = {}
        self.episode_data = {name: torch.zeros((N,)) for name in spec["rewards"]}

        # --- command_manager.py:66-81, 127-130
        self.command = {}
        for name, c in spec["commands"].items():
            if c.get("type") == "gait":  # the gait_trainer example's own manager (oracle/gait_port.py)
                from .gait_port import GaitPort

                gait = GaitPort(self, c)
                gait.build()
                self.command[name] = {"gait": gait, "python": None, "command": gait._command}
                setattr(self, name, gait)
                continue
            rng = c["range"]
            k = len(rng) if isinstance(rng, dict) else 1
            self.command[name] = {
                "range": rng, "command": torch.zeros(N, k),
                "resample_steps": int(c["resample_time_sec"] / self.dt),
                "python": c if c.get("type") == "python" else None,
            }
            setattr(self, name, _CommandView(self.command[name]))

        # --- entity_manager.py:87-99, 157-167
        self.global_gravity = torch.tensor([0.0, 0.0, -1.0]).repeat(N, 1)
        self.base_pos = torch.zeros((N, 3))
        self.base_quat = torch.zeros((N, 4))
        self.inv_base_quat = torch.zeros_like(self.base_quat)
        self._entity_cache()
        self.reset_items = []
        for name, item in spec["entity"]["on_reset"].items():
            p = dict(item.get("params") or {})
            entry = {"fn": item["fn"], "params": p}
            if item["fn"] == "position":  # mdp/reset.py:81-100
                entry["reset_pos"] = torch.tensor(p["position"])
                entry["pos_buffer"] = torch.zeros((N, 3))
                entry["reset_quat"] = torch.tensor(p["quat"]) if p.get("quat") is not None else None
                entry["quat_buffer"] = torch.zeros((N, 4)) if p.get("quat") is not None else None
                entry["zero_velocity"] = p.get("zero_velocity", True)
            elif item["fn"] == "randomize_terrain_position":  # mdp/reset.py:146-170
                entry["rotation"] = p.get("rotation", {"z": (0, 2 * math.pi)})
                entry["rotation_buffer"] = torch.zeros((N, 3))
                entry["quat_buffer"] = torch.zeros((N, 4))
            self.reset_items.append(entry)

        # --- observation_manager.py:180-216 (dry run sizes the space; consumes RNG when noisy)
        self.obs_groups = {}
        for group, g in spec["observations"].items():
            og = {"terms": g["terms"], "noise": g.get("noise"), "history_len": g.get("history_len") or 1}
            self.obs_groups[group] = og
            obs = self._perform_observation(group)
            og["single"] = obs.shape[1]
            og["history"] = [torch.zeros((N, og["single"])) for _ in range(og["history_len"])]

    # ------------------------------------------------------------------------------------------
    # entity helpers (entity_manager.py:130-146, 189-195; utils.py:13-55)
    # ------------------------------------------------------------------------------------------
    def _entity_cache(self):
        self.base_pos[:] = self.robot.get_pos()
        self.base_quat[:] = self.robot.get_quat()
        self.inv_base_quat = inv_quat(self.base_quat)

    def _lin_vel(self, cached=True):
        if cached:
            return transform_by_quat(self.robot.get_vel(), self.inv_base_quat)
        return transform_by_quat(self.robot.get_vel(), inv_quat(self.robot.get_quat()))

    def _ang_vel(self, cached=True):
        if cached:
            return transform_by_quat(self.robot.get_ang(), self.inv_base_quat)
        return transform_by_quat(self.robot.get_ang(), inv_quat(self.robot.get_quat()))

    def _gravity(self, cached=True):
        if cached:
            return transform_by_quat(self.global_gravity, self.inv_base_quat)
        q = inv_quat(self.robot.get_quat())
        g = torch.tensor([0.0, 0.0, -1.0]).expand(q.shape[0], 3)
        return transform_by_quat(g, q)

    @staticmethod
    def _cached(params) -> bool:
        return params.get("entity_manager") is not None

    # ------------------------------------------------------------------------------------------
    # terrain (terrain_manager.py:92-279)
    # ------------------------------------------------------------------------------------------
    def _terrain_height(self, x, y):
        n = x.shape[0]
        (x_min, x_max, y_min, y_max) = self.t_bounds
        norm_x = self.t_norm[:n, 0]
        norm_y = self.t_norm[:n, 1]
        norm_x.copy_(x)
        norm_x.sub_(x_min)
        norm_x.div_(x_max - x_min)
        norm_x.mul_(2)
        norm_x.sub_(1)
        norm_y.copy_(y)
        norm_y.sub_(y_min)
        norm_y.div_(y_max - y_min)
        norm_y.mul_(2)
        norm_y.sub_(1)
        grid = self.t_grid[:n]
        grid[:, 0, 0, 0] = norm_x
        grid[:, 0, 0, 1] = norm_y
        interpolated = F.grid_sample(
            self.t_height_field.unsqueeze(0).expand(n, -1, -1, -1), grid,
            mode="bilinear", padding_mode="border", align_corners=True,
        )
        heights = self.t_heights[:n]
        heights.copy_(interpolated[:, 0, 0, 0])
        return heights

    def _random_env_pos(self, envs_idx, usable_ratio=0.5, height_offset=0.1e-3):
        output = self.t_env_pos
        (x_origin, _, y_origin, _) = self.t_bounds
        (x_size, y_size) = self.t_size
        usable_x, usable_y = x_size * usable_ratio, y_size * usable_ratio
        buf_x, buf_y = (x_size - usable_x) / 2, (y_size - usable_y) / 2
        x_min, x_max = x_origin + buf_x, x_origin + x_size - buf_x
        y_min, y_max = y_origin + buf_y, y_origin + y_size - buf_y
        rx = torch.rand_like(output[envs_idx, 0])
        self.rng_log.append(("spawn_x", rx.clone()))
        output[envs_idx, 0] = rx * (x_max - x_min) + x_min
        ry = torch.rand_like(output[envs_idx, 1])
        self.rng_log.append(("spawn_y", ry.clone()))
        output[envs_idx, 1] = ry * (y_max - y_min) + y_min
        heights = self._terrain_height(output[envs_idx, 0], output[envs_idx, 1])
        output[envs_idx, 2] = heights + height_offset
        return output[envs_idx]

    # ------------------------------------------------------------------------------------------
    # step  (managed_env.py:286-334)
    # ------------------------------------------------------------------------------------------
    def step(self, actions: torch.Tensor):
        N = self.num_envs
        self.rng_log = []
        self.margins = []
        # -- genesis_env.py:193-203
        self.extras = {"episode": {}}
        self.step_count += 1
        self.episode_length += 1
        if self.actions is None:
            self.actions = actions.detach().clone()
            self.last_actions = torch.zeros_like(actions)
        else:
            self.last_actions[:] = self.actions[:]
            self.actions[:] = actions[:]
        self.extras["observations"] = {}

        # -- action manager: base.py:67-82, position_action_manager.py:376-419
        if "action" not in self.disabled:
            a = actions
            if self.delay_step > 0:
                self.action_delay_buffer.insert(0, a)
                a = self.action_delay_buffer.pop()
            self.raw_actions = a
            if self.targets is None:
                self.targets = a.clone()
            else:
                self.targets[:] = a[:]
            self.targets = self._handle_actions(self.targets)

        self.scene.step()

        self._entity_cache()                                  # entity_manager.py:185-195
        for m in self.contact.values():                       # contact_manager.py:331-336
            self._contact_forces(m)
            self._air_time(m)

        terminated, truncated = self._terminations()          # termination_manager.py:151-190
        reset_idx = (terminated | truncated).nonzero(as_tuple=False).reshape((-1,)).detach()
        self.reset_idx = reset_idx

        rewards = self._rewards()                             # reward_manager.py:166-195

        self.resample_idx = {}
        for name, c in self.command.items():                  # command_manager.py:152-162
            if c.get("gait") is not None:
                c["gait"].step()
                continue
            if c["python"] is not None:  # user-level subclass overriding step()
                c["python"]["step"](getattr(self, name), self)
                continue
            idx = (self.episode_length % c["resample_steps"] == 0).nonzero(as_tuple=False).reshape((-1,))
            self.resample_idx[name] = idx
            self._resample(name, c, idx, "cmd_step")

        if reset_idx.numel() > 0:
            self.reset(reset_idx)

        obs = self._get_observations()
        return obs, rewards, terminated, truncated, self.extras

    def _handle_actions(self, actions):
        if self.within_limits:  # position_within_limits.py:113-131
            actions.clamp_(-1.0, 1.0)
            out = actions * self.wl_scale + self.wl_offset
            self.robot.control_dofs_position(out, self.dofs_idx)
            return out
        if torch.isnan(actions).any():
            self.printed.append("nan_actions")
        if torch.isinf(actions).any():
            self.printed.append("inf_actions")
        actions = actions * self.scale_values + self.offset_values
        actions = torch.clamp(actions, min=self.clip_values[:, 0], max=self.clip_values[:, 1])
        self.robot.control_dofs_position(actions, self.dofs_idx)
        return actions

    # -- contacts ------------------------------------------------------------------------------
    def _contact_forces(self, m):
        """contact_manager.py:384-432"""
        c = self.scene.rigid_solver.collider.get_contacts(as_tensor=True, to_torch=True)
        force, link_a, link_b, position = c["force"], c["link_a"], c["link_b"], c["position"]
        if torch.isnan(force).any() or torch.isinf(force).any():
            force = torch.nan_to_num(force, nan=0.0, posinf=0.0, neginf=0.0)
            self.printed.append("contact_nan")
        links_quat = self.scene.rigid_solver.get_links_quat()
        m["contacts"].fill_(0.0)
        m["positions"].fill_(0.0)
        m["counts"].fill_(0.0)
        kernel_get_contact_forces(
            force.contiguous(), position.contiguous(), link_a.contiguous(), link_b.contiguous(),
            links_quat.contiguous(), m["link_ids"].contiguous(), m["with_link_ids"].contiguous(),
            m["contacts"], m["positions"], m["counts"], 1 if m["has_filter"] else 0,
        )

    def _air_time(self, m):
        """contact_manager.py:434-477"""
        if not m["track"]:
            return
        dt = self.scene.dt
        norm = torch.norm(m["contacts"][:, :, :], dim=-1)
        self._margin("air_time_contact", norm, m["threshold"])
        is_contact = norm > m["threshold"]
        is_new_contact = (m["cur_air"] > 0) * is_contact
        is_new_detached = (m["cur_contact"] > 0) * ~is_contact
        m["last_air"] = torch.where(is_new_contact, m["cur_air"] + dt, m["last_air"])
        m["cur_air"] = torch.where(~is_contact, m["cur_air"] + dt, 0.0)
        m["last_contact"] = torch.where(is_new_detached, m["cur_contact"] + dt, m["last_contact"])
        m["cur_contact"] = torch.where(is_contact, m["cur_contact"] + dt, 0.0)

    # -- terminations (mdp/terminations.py) ------------------------------------------------------
    def _termination_value(self, fn, p: dict) -> torch.Tensor:
        if callable(fn):  # user-defined term
            return fn(self, **p)
        if fn == "timeout":  # :17-23
            if self.max_episode_length is None:
                return torch.zeros(self.num_envs, dtype=torch.bool)
            return self.episode_length > self.max_episode_length
        if fn == "bad_orientation":  # :26-71
            in_grace = self.episode_length <= p.get("grace_steps", 0)
            g = self._gravity(self._cached(p))
            tilt = torch.norm(g[:, :2], dim=1)
            angle = torch.asin(torch.clamp(tilt, max=0.99))
            limit = math.radians(p.get("limit_angle", 40.0))
            self._margin("bad_orientation_angle", angle, limit)
            return (~in_grace) & (angle > limit)
        if fn == "base_height_below_minimum":  # :74-99
            base_pos = self.base_pos if self._cached(p) else self.robot.get_pos()
            return base_pos[:, 2] < p.get("minimum_height", 0.05)
        if fn == "out_of_bounds":  # :102-137
            position = self.robot.get_pos()
            (x_min, x_max, y_min, y_max) = self.t_bounds
            margin = p.get("border_margin", 0.5)
            x_lo, x_hi = x_min + margin, x_max - margin
            y_lo, y_hi = y_min + margin, y_max - margin
            x_pos, y_pos = position[:, 0], position[:, 1]
            return (x_pos < x_lo) | (x_pos > x_hi) | (y_pos < y_lo) | (y_pos > y_hi)
        m = self.contact[p["contact_manager"][1:]] if "contact_manager" in p else None
        if fn == "has_contact":  # :139-155
            norm = m["contacts"][:, :].norm(dim=-1)
            self._margin("term_has_contact", norm, p.get("threshold", 1.0))
            has = norm > p.get("threshold", 1.0)
            return has.sum(dim=1) >= p.get("min_contacts", 1)
        if fn == "contact_force":  # :158-172
            norm = torch.norm(m["contacts"], dim=-1)
            self._margin("term_contact_force", norm, p.get("threshold", 1.0))
            return torch.any(norm > p.get("threshold", 1.0), dim=-1)
        if fn == "contact_force_with_grace_period":  # :175-205
            in_grace = self.episode_length <= p.get("grace_steps", 10)
            norm = torch.norm(m["contacts"], dim=-1)
            self._margin("term_contact_force_grace", norm, p.get("threshold", 100.0))
            exceeded = torch.any(norm > p.get("threshold", 100.0), dim=-1)
            return (~in_grace) & exceeded.detach()
        raise KeyError(fn)

    def _terminations(self):
        if "termination" in self.disabled:
            return self.terminated, self.truncated
        self.terminated[:] = False
        self.truncated[:] = False
        logging = self.extras["episode"]
        for name, item in self.spec["terminations"].items():
            value = self._termination_value(item["fn"], item.get("params") or {})
            if item.get("time_out", False):
                self.truncated |= value
            else:
                self.terminated |= value
            dones = value.nonzero(as_tuple=True)[0]
            if dones.numel() > 0:
                logging[f"Terminations / {name}"] = value.float().mean().detach()
        self.extras["terminations"] = self.terminated
        self.extras["time_outs"] = self.truncated
        return self.terminated, self.truncated

    # -- rewards (mdp/rewards.py) ----------------------------------------------------------------
    def _resolve_cmd(self, ref: str) -> torch.Tensor:
        return self.command[ref[1:]]["command"]

    def _reward_value(self, fn, p: dict, name: str = "") -> torch.Tensor:
        if callable(fn):  # user-defined term
            return fn(self, **p)
        if fn.startswith("@"):  # a term that is a method of the gait command manager
            owner, method = fn[1:].split(".")
            gait = self.command[owner]["gait"]
            if method == "gait_phase_reward":
                return gait.gait_phase_reward(self.contact[p["contact_manager"][1:]])
            return gait.foot_height_reward(**p)
        if fn == "is_alive":  # :31-37
            return (~self.extras["terminations"]).float().detach()
        if fn == "terminated":  # :40-46
            return self.extras["terminations"].float().detach()
        if fn == "base_height":  # :54-90
            base_pos = self.robot.get_pos()
            height_offset = 0.0
            if p.get("terrain_manager") is not None:
                height_offset = self._terrain_height(base_pos[:, 0], base_pos[:, 1])
            target = p.get("target_height")
            if p.get("height_command") is not None:
                target = self._resolve_cmd(p["height_command"]).squeeze(-1)
            return torch.square(base_pos[:, 2] - height_offset - target)
        if fn == "dof_similar_to_default":  # :93-109
            dof_pos = self.robot.get_dofs_position(self.dofs_idx)
            return torch.sum(torch.abs(dof_pos - self.default_dofs_pos), dim=1)
        if fn == "lin_vel_z_l2":  # :112-135
            return torch.square(self._lin_vel(self._cached(p))[:, 2])
        if fn == "ang_vel_xy_l2":  # :138-161
            return torch.sum(torch.square(self._ang_vel(self._cached(p))[:, :2]), dim=1)
        if fn == "flat_orientation_l2":  # :164-193
            return torch.sum(torch.square(self._gravity(self._cached(p))[:, :2]), dim=1)
        if fn == "body_acceleration_exp":  # :196-249 (class-style term with state, not cleared on reset)
            state = self.body_acc_state.setdefault(name, {})
            curr_lin = self._lin_vel(True)
            curr_ang = self._ang_vel(True)
            if "prev_lin" in state:
                lin_acc = (curr_lin - state["prev_lin"]) / self.dt
                ang_acc = (curr_ang - state["prev_ang"]) / self.dt
            else:
                lin_acc = torch.zeros_like(curr_lin)
                ang_acc = torch.zeros_like(curr_ang)
            state["prev_lin"] = curr_lin.clone()
            state["prev_ang"] = curr_ang.clone()
            pelvis_motion = torch.norm(lin_acc, dim=-1) + torch.norm(ang_acc, dim=-1)
            return 1 - torch.exp(-p.get("sensitivity", 0.10) * pelvis_motion)
        if fn == "action_rate_l2":  # :257-271
            return torch.sum(torch.square(self.last_actions - self.actions), dim=1)
        if fn == "command_tracking_lin_vel":  # :279-317
            v = self._lin_vel(self._cached(p))
            if p.get("vel_cmd_manager") is not None:
                command = self._resolve_cmd(p["vel_cmd_manager"])[:, :2]
            else:
                command = self.fixed_command[:, :2]
            err = torch.sum(torch.square(command - v[:, :2]), dim=1)
            return torch.exp(-err / p.get("sensitivity", 0.25))
        if fn == "command_tracking_ang_vel":  # :320-358
            w = self._ang_vel(self._cached(p))
            if p.get("vel_cmd_manager") is not None:
                commanded = self._resolve_cmd(p["vel_cmd_manager"])[:, 2]
            else:
                commanded = self.fixed_command[:, 2]
            err = torch.square(commanded - w[:, 2])
            return torch.exp(-err / p.get("sensitivity", 0.25))
        if fn == "stand_still_joint_deviation_l1":  # :361-385
            command = self._resolve_cmd(p["vel_cmd_manager"])
            joint_pos = self.robot.get_dofs_position(self.dofs_idx)
            deviation = torch.sum(torch.abs(joint_pos - self.default_dofs_pos), dim=1)
            cmd_norm = torch.norm(command[:, :2], dim=1)
            self._margin("stand_still_cmd", cmd_norm, p.get("command_threshold", 0.06))
            return deviation * (cmd_norm < p.get("command_threshold", 0.06))
        m = self.contact[p["contact_manager"][1:]] if "contact_manager" in p else None
        if fn == "has_contact":  # :393-410
            norm = m["contacts"][:, :].norm(dim=-1)
            self._margin("rew_has_contact", norm, p.get("threshold", 1.0))
            has = norm > p.get("threshold", 1.0)
            return (has.sum(dim=1) >= p.get("min_contacts", 1)).float()
        if fn == "contact_force":  # :413-428
            violation = torch.norm(m["contacts"][:, :, :], dim=-1) - p.get("threshold", 1.0)
            return torch.sum(violation.clip(min=0.0), dim=1)
        if fn == "feet_air_time":  # :431-469 with contact_manager.py:198-224
            in_contact = m["cur_contact"] > 0.0
            recent = m["cur_contact"] < (self.dt + 1.0e-8)
            made_contact = in_contact * recent
            air_time = (m["last_air"] - p["time_threshold"]) * made_contact
            if p.get("time_threshold_max") is not None:
                air_time = torch.clamp(air_time, max=p["time_threshold_max"] - p["time_threshold"])
            reward = torch.sum(air_time, dim=1)
            if p.get("vel_cmd_manager") is not None:
                cmd_norm = torch.norm(self._resolve_cmd(p["vel_cmd_manager"])[:, :2], dim=1)
                self._margin("feet_air_cmd", cmd_norm, 0.1)
                reward *= cmd_norm > 0.1
            return reward
        if fn == "feet_slide":  # :472-504
            norm = torch.norm(m["contacts"][:, :, :], dim=-1)
            self._margin("feet_slide_contact", norm, 1.0)
            in_contact = norm > 1.0
            link_vel = self.robot.get_links_vel(links_idx_local=m["local_link_ids"])
            return torch.sum(link_vel.norm(dim=-1) * in_contact, dim=1)
        raise KeyError(fn)

    def _rewards(self):
        if "reward" in self.disabled:
            return self.reward_buf
        dt = self.dt
        self.reward_buf[:] = 0.0
        self.episode_seconds += dt
        for name, item in self.spec["rewards"].items():
            if item["weight"] == 0:
                continue
            weight = item["weight"] * dt
            value = self._reward_value(item["fn"], item.get("params") or {}, name) * weight
            self.reward_buf += value
            self.episode_data[name] += value
        return self.reward_buf

    # -- commands (command_manager.py:290-303) ---------------------------------------------------
    def _resample(self, name, c, env_ids, tag):
        rng = c["range"]
        ranges = list(rng.values()) if isinstance(rng, dict) else [rng]
        buffer = torch.empty(len(env_ids))
        for i in range(c["command"].shape[1]):
            buffer.uniform_(*ranges[i])
            self.rng_log.append((f"{tag}:{name}:{i}", buffer.clone()))
            c["command"][env_ids, i] = buffer

    # ------------------------------------------------------------------------------------------
    # reset  (managed_env.py:336-371)
    # ------------------------------------------------------------------------------------------
    def reset(self, env_ids=None):
        N = self.num_envs
        initial = env_ids is None
        if initial:
            self.rng_log = []
        # -- genesis_env.py:221-254
        idx = torch.arange(N) if env_ids is None else env_ids
        if self.step_count == 0:
            self.actions = torch.zeros((N, self.num_actions), dtype=torch.float32)
            self.last_actions = torch.zeros_like(self.actions)
        if idx.numel() > 0:
            if self.actions is not None:
                self.actions[idx] = 0.0
                self.last_actions[idx] = 0.0
            self.episode_length[idx] = 0
        if len(idx) > 0 and self.max_episode_random_scaling > 0.0 and self.base_max_episode_length is not None:
            max_random_scaling = self.base_max_episode_length * self.max_episode_random_scaling
            u = self._draw_max_len(idx)
            randomization = u * max_random_scaling
            self.max_episode_length[idx] = torch.round(self.base_max_episode_length + randomization).to(torch.int32)

        # -- position_action_manager.py:421-464 (envs_idx=None -> arange)
        aidx = torch.arange(N) if env_ids is None else env_ids
        ns = self.action_noise_scale

        def noisy(tag, values):
            if ns == 0.0:
                return values
            u = torch.empty_like(values).uniform_(-1, 1)
            self.rng_log.append((f"action_dr:{tag}", u.clone()))
            return values + u * ns

        setters = {"pd_kp": "set_dofs_kp", "pd_kv": "set_dofs_kv", "damping": "set_dofs_damping",
                   "stiffness": "set_dofs_stiffness", "frictionloss": "set_dofs_frictionloss"}
        if "action" not in self.disabled:  # position_action_manager.py:426-427
            for key, setter in setters.items():
                if key in self.gain_values:
                    getattr(self.robot, setter)(noisy(key, self.gain_values[key]), self.dofs_idx, aidx)
            if self.force_range is not None:
                lower = noisy("force_lower", self.force_range[0])
                upper = noisy("force_upper", self.force_range[1])
                self.robot.set_dofs_force_range(lower, upper, self.dofs_idx, aidx)
            position = noisy("position", self.default_dofs_pos[aidx])
            self.robot.set_dofs_position(position=position, dofs_idx_local=self.dofs_idx, envs_idx=aidx)

        # -- entity_manager.py:169-183 + mdp/reset.py
        eidx = torch.arange(N) if env_ids is None else env_ids
        for item in self.reset_items:
            if item["fn"] == "position":  # reset.py:102-124
                item["pos_buffer"][eidx] = item["reset_pos"]
                self.robot.set_pos(item["pos_buffer"][eidx], envs_idx=eidx, zero_velocity=item["zero_velocity"])
                if item["reset_quat"] is not None:
                    item["quat_buffer"][eidx] = item["reset_quat"].reshape(1, -1)
                    self.robot.set_quat(item["quat_buffer"][eidx], envs_idx=eidx, zero_velocity=item["zero_velocity"])
            elif item["fn"] == "randomize_terrain_position":  # reset.py:172-226
                p = item["params"]
                pos = self._random_env_pos(eidx, height_offset=p.get("height_offset", 0.1e-3))
                zero_velocity = p.get("zero_velocity", True)
                self.robot.set_pos(pos, envs_idx=eidx, zero_velocity=zero_velocity)
                rotation = item["rotation"]
                if rotation is not None:
                    for axis, col in (("x", 0), ("y", 1), ("z", 2)):
                        v = rotation.get(axis, 0)
                        if isinstance(v, tuple):
                            item["rotation_buffer"][eidx, col] = self._uniform(f"spawn_rot_{axis}", len(eidx), *v)
                    item["quat_buffer"][eidx] = xyz_to_quat(item["rotation_buffer"][eidx])
                    self.robot.set_quat(item["quat_buffer"][eidx], envs_idx=eidx, zero_velocity=zero_velocity)
            elif item["fn"] == "zero_all_dofs_velocity":
                self.robot.zero_all_dofs_velocity(eidx)
            else:
                raise KeyError(item["fn"])

        # -- contact_manager.py:316-329
        cidx = torch.arange(N) if env_ids is None else env_ids
        for m in self.contact.values():
            if m["track"]:
                m["cur_air"][cidx] = 0.0
                m["cur_contact"][cidx] = 0.0
                m["last_air"][cidx] = 0.0
                m["last_contact"][cidx] = 0.0

        # -- reward_manager.py:197-222
        ridx = torch.arange(N) if env_ids is None else env_ids
        logging = self.extras["episode"]
        if "reward" not in self.disabled:
            episode_seconds = self.episode_seconds[ridx]
            for name, value in self.episode_data.items():
                if self.spec["rewards"][name]["weight"] != 0:
                    value[ridx] /= episode_seconds
                    episode_mean = torch.mean(value[ridx])
                    self.episode_mean[name] = episode_mean.item()
                    logging[f"Rewards / {name}"] = episode_mean
                self.episode_data[name][ridx] = 0.0
        self.episode_seconds[ridx] = 1e-10

        # -- command_manager.py:164-170
        for name, c in self.command.items():
            if c.get("gait") is not None:
                c["gait"].reset(env_ids)
                continue
            if c["python"] is not None:
                c["python"]["reset"](getattr(self, name), self, env_ids)
                continue
            kidx = torch.arange(N) if env_ids is None else env_ids
            self._resample(name, c, kidx, "cmd_reset")

        obs = None
        if initial:
            obs = self._get_observations()
        return obs, self.extras

    # ------------------------------------------------------------------------------------------
    # observations (observation_manager.py:218-256, managed_env.py:373-396)
    # ------------------------------------------------------------------------------------------
    def _obs_value(self, term: dict) -> torch.Tensor:
        kind = term["fn"]
        if callable(kind):  # user-defined term
            return kind(env=self)
        if kind == "ang_vel_uncached":
            return self._ang_vel(False)
        if kind == "command":
            c = self.command[term["mgr"]]
            return c["gait"].observation() if c.get("gait") is not None else c["command"]
        if kind == "ang_vel":
            return self._ang_vel(True)
        if kind == "lin_vel":
            return self._lin_vel(True)
        if kind == "gravity":
            return self._gravity(True)
        if kind == "dof_pos":
            return self.robot.get_dofs_position(self.dofs_idx)
        if kind == "dof_vel":
            return self.robot.get_dofs_velocity(self.dofs_idx)
        if kind in ("dof_force", "entity_dofs_force"):  # mdp/observations.py:133-158 with an action manager
            return self.robot.get_dofs_force(self.dofs_idx)
        if kind in ("actions", "current_actions"):  # base.py:96-102
            if self.targets is None:
                return torch.zeros((self.num_envs, self.num_actions))
            return self.targets
        if kind == "contact_force":  # mdp/observations.py:181-193
            return torch.norm(self.contact[term["mgr"]]["contacts"][:, :, :], dim=-1)
        raise KeyError(kind)

    def _perform_observation(self, group: str) -> torch.Tensor:
        og = self.obs_groups[group]
        obs = []
        for name, term in og["terms"].items():
            value = self._obs_value(term)
            scale = term.get("scale", 1.0)
            if scale is not None and scale != 1.0:
                value *= scale
            noise = term.get("noise", None) or og["noise"]
            if noise is not None and noise != 0.0:
                u = torch.empty_like(value).uniform_(-1, 1)
                self.rng_log.append((f"obs_noise:{group}:{name}", u.clone()))
                value += u * noise
            obs.append(value)
        return torch.cat(obs, dim=-1)

    def _get_observations(self):
        if "observations" not in self.extras:
            self.extras["observations"] = {}
        if "policy" in self.extras["observations"]:
            return self.extras["observations"]["policy"]
        policy = None
        for group, og in self.obs_groups.items():
            og["history"].pop()
            og["history"].insert(0, self._perform_observation(group))
            obs = torch.cat(og["history"], dim=-1)
            self.extras["observations"][group] = obs
            if group == "policy":
                policy = obs
        return policy

    # ------------------------------------------------------------------------------------------
    # snapshot of every buffer (used for bit-exact comparison against the reference and the kernels)
    # ------------------------------------------------------------------------------------------
    def snapshot(self) -> dict[str, torch.Tensor]:
        s = {
            "episode_length": self.episode_length, "max_episode_length": self.max_episode_length,
            "actions": self.actions, "last_actions": self.last_actions,
            "targets": self.targets if self.targets is not None else torch.zeros((self.num_envs, self.num_actions)),
            "base_pos": self.base_pos, "base_quat": self.base_quat, "inv_base_quat": self.inv_base_quat,
            "terminated": self.terminated, "truncated": self.truncated,
            "reward_buf": self.reward_buf, "episode_seconds": self.episode_seconds,
        }
        for name, v in self.episode_data.items():
            s[f"episode_data/{name}"] = v
        for name, c in self.command.items():
            s[f"command/{name}"] = c["command"]
            if c.get("gait") is not None:
                s.update(c["gait"].snapshot(name))
        for name, m in self.contact.items():
            s[f"contact/{name}/contacts"] = m["contacts"]
            s[f"contact/{name}/positions"] = m["positions"]
            if m["track"]:
                for k in ("last_air", "cur_air", "last_contact", "cur_contact"):
                    s[f"contact/{name}/{k}"] = m[k]
        return {k: v.clone() for k, v in s.items() if v is not None}
