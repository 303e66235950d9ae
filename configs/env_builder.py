"""
Builds a ManagedEnvironment from a term-table spec using a given
*namespace* of manager classes and mdp functions.

Because the drop-in (genesis_forge_b200) mirrors the reference's manager API, the very same builder
produces
  * the UNMODIFIED reference environment   (namespace = reference_namespace(), needs /root/reference)
  * the CUDA drop-in environment           (namespace = dropin_namespace())
which is itself the drop-in claim under test: identical `config()` code, two implementations.

The `config()` body below is written the way the reference examples write theirs (e.g.
examples/command_direction/environment.py:85-247): EntityManager with on_reset items, a
PositionActionManager, command / contact managers, RewardManager, TerminationManager and an
ObservationManager whose terms are lambdas around manager getters.
"""
from __future__ import annotations

import types

import torch

from genesis_forge_b200.synthetic import ROBOT_MODELS, SyntheticScene


def reference_namespace():
    """Manager classes / mdp modules of the unmodified reference (imports it under the shim)."""
    from oracle import shim

    gf = shim.import_reference()
    import genesis_forge.managers as managers
    import genesis_forge.mdp as mdp

    return types.SimpleNamespace(
        name="reference",
        ManagedEnvironment=gf.ManagedEnvironment,
        managers=managers,
        rewards=mdp.rewards,
        terminations=mdp.terminations,
        observations=mdp.observations,
        reset=mdp.reset,
    )


def dropin_namespace():
    """Manager classes / mdp modules of the B200 drop-in."""
    import genesis_forge_b200 as gfb
    import genesis_forge_b200.managers as managers
    import genesis_forge_b200.mdp as mdp

    return types.SimpleNamespace(
        name="dropin",
        ManagedEnvironment=gfb.ManagedEnvironment,
        managers=managers,
        rewards=mdp.rewards,
        terminations=mdp.terminations,
        observations=mdp.observations,
        reset=mdp.reset,
    )


def make_scene(spec: dict, device, source=None, copy_on_get=False, n_contacts=8, seed=1234, pool=0,
               apply_setters=True):
    scene = SyntheticScene(
        dt=spec["dt"], n_contacts=n_contacts, seed=seed, device=device, source=source,
        copy_on_get=copy_on_get, pool=pool, apply_setters=apply_setters,
        source_kw={"xy_range": spec["xy_range"]} if "xy_range" in spec else None,
    )
    if "terrain" in spec:
        terrain = scene.add_terrain(**spec["terrain"])
    else:
        terrain = scene.add_plane()
    robot = scene.add_robot(ROBOT_MODELS[spec["robot"]])
    return scene, terrain, robot


def build_env(spec: dict, ns, num_envs: int, device, **scene_kw):
    """Instantiate (not yet build()) an environment of namespace `ns` for `spec`."""

    class SpecEnv(ns.ManagedEnvironment):
        def __init__(self):
            super().__init__(
                num_envs=num_envs,
                dt=spec["dt"],
                max_episode_length_sec=spec.get("max_episode_length_sec", 10),
                max_episode_random_scaling=spec.get("max_episode_random_scaling", 0.0),
            )
            self.scene, self.terrain, self.robot = make_scene(spec, device, **scene_kw)
            if "fixed_command" in spec:
                self.fixed_command = torch.zeros((num_envs, 3), device=device, dtype=torch.float32)
                for i, v in enumerate(spec["fixed_command"]):
                    self.fixed_command[:, i] = v

        # ---- helpers --------------------------------------------------------------------------
        def managers_by_name(self, kind):
            """Attribute names of the command / contact managers, in creation order."""
            return list(spec["commands" if kind == "command" else "contacts"].keys())

        def _resolve(self, value):
            if isinstance(value, str) and value.startswith("@"):
                ref = value[1:]
                if ref == "fixed_command[:, :2]":
                    return self.fixed_command[:, :2]
                if ref == "fixed_command[:, 2]":
                    return self.fixed_command[:, 2]
                return getattr(self, ref)
            if torch.is_tensor(value):  # per-env tensors given in a spec live on the env's device
                return value.to(device)
            return value

        def _params(self, params):
            return {k: self._resolve(v) for k, v in (params or {}).items()}

        def _term_fn(self, fn, module):
            """A term function: a callable, "@<manager>.<method>" (a manager's own term) or an mdp name."""
            if callable(fn):
                return fn
            if fn.startswith("@"):
                owner, method = fn[1:].split(".")
                return getattr(getattr(self, owner), method)
            return getattr(module, fn)

        def _obs_fn(self, term):
            kind = term["fn"]
            if callable(kind):  # user-defined observation term
                return kind, {}
            if kind == "ang_vel_uncached":  # mdp.observations without an entity manager (utils path)
                return ns.observations.entity_angular_velocity, {}
            if kind == "command":
                return getattr(self, term["mgr"]).observation, {}
            if kind == "ang_vel":
                return (lambda env: self.robot_manager.get_angular_velocity()), {}
            if kind == "lin_vel":
                return (lambda env: self.robot_manager.get_linear_velocity()), {}
            if kind == "gravity":
                return (lambda env: self.robot_manager.get_projected_gravity()), {}
            if kind == "dof_pos":
                return (lambda env: self.action_manager.get_dofs_position()), {}
            if kind == "dof_vel":
                return (lambda env: self.action_manager.get_dofs_velocity()), {}
            if kind == "dof_force":
                return (lambda env: self.action_manager.get_dofs_force()), {}
            if kind == "entity_dofs_force":
                return ns.observations.entity_dofs_force, {"action_manager": self.action_manager}
            if kind == "actions":
                return (lambda env: self.action_manager.get_actions()), {}
            if kind == "current_actions":
                return ns.observations.current_actions, {"action_manager": self.action_manager}
            if kind == "contact_force":
                return ns.observations.contact_force, {"contact_manager": getattr(self, term["mgr"])}
            raise KeyError(kind)

        # ---- the manager table ----------------------------------------------------------------
        def config(self):
            M = ns.managers
            if "terrain" in spec:
                self.terrain_manager = M.TerrainManager(self)

            on_reset = {}
            for name, item in spec["entity"]["on_reset"].items():
                on_reset[name] = {"fn": getattr(ns.reset, item["fn"]), "params": self._params(item.get("params"))}
            self.robot_manager = M.EntityManager(self, entity_attr="robot", on_reset=on_reset)
            if hasattr(self, "after_entity_manager"):  # workloads with further entities (configs/second_entity.py)
                self.after_entity_manager()

            a = dict(spec["action"])
            kind = a.pop("type")
            cls = M.PositionActionManager if kind == "position" else M.PositionWithinLimitsActionManager
            self.action_manager = cls(self, **a)

            for name, c in spec["commands"].items():
                c = dict(c)
                ctype = c.pop("type")
                if ctype == "python":  # user-level subclass with its own step()/reset()
                    env_self, step_fn, reset_fn = self, c["step"], c["reset"]

                    class PythonCommand(M.CommandManager):
                        def step(mgr, _fn=step_fn):
                            _fn(mgr, env_self)

                        def reset(mgr, env_ids=None, _fn=reset_fn):
                            _fn(mgr, env_self, env_ids)

                    mgr = PythonCommand(self, range=c["range"], resample_time_sec=c["resample_time_sec"])
                elif ctype == "gait":  # the example's user-level manager (examples/gait_trainer)
                    from configs import gait

                    if ns.name == "reference":
                        cls = gait.reference_gait_command_manager()  # the unmodified example class
                    else:
                        cls = gait.make_gait_command_manager(M.CommandManager)
                    mgr = cls(self, foot_names=c["foot_names"], resample_time_sec=c["resample_time_sec"])
                    for what, times in (c.get("curriculum") or {}).items():
                        for _ in range(times):
                            getattr(mgr, f"increment_{what}")()
                elif ctype == "velocity":
                    mgr = M.VelocityCommandManager(
                        self, range=c["range"], resample_time_sec=c["resample_time_sec"],
                        standing_probability=c.get("standing_probability", 0.0),
                    )
                else:
                    mgr = M.CommandManager(self, range=c["range"], resample_time_sec=c["resample_time_sec"])
                setattr(self, name, mgr)

            for name, c in spec["contacts"].items():
                setattr(self, name, M.ContactManager(self, **c))

            cfg = {}
            for name, item in spec["rewards"].items():
                cfg[name] = {
                    "weight": item["weight"],
                    "fn": self._term_fn(item["fn"], ns.rewards),
                    "params": self._params(item.get("params")),
                }
            self.reward_manager = M.RewardManager(self, logging_enabled=True, cfg=cfg)

            term_cfg = {}
            for name, item in spec["terminations"].items():
                term_cfg[name] = {
                    "fn": self._term_fn(item["fn"], ns.terminations),
                    "time_out": item.get("time_out", False),
                    "params": self._params(item.get("params")),
                }
            self.termination_manager = M.TerminationManager(self, logging_enabled=True, term_cfg=term_cfg)

            self.observation_managers = {}
            for group, g in spec["observations"].items():
                ocfg = {}
                for name, term in g["terms"].items():
                    fn, params = self._obs_fn(term)
                    entry = {"fn": fn}
                    if params:
                        entry["params"] = params
                    if "scale" in term:
                        entry["scale"] = term["scale"]
                    if "noise" in term:
                        entry["noise"] = term["noise"]
                    ocfg[name] = entry
                self.observation_managers[group] = M.ObservationManager(
                    self, cfg=ocfg, name=group, history_len=g.get("history_len"), noise=g.get("noise"),
                )

    return SpecEnv()
