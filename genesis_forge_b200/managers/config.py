"""
Term-table plumbing: the dict entries users write in manager configs are wrapped in config items.

Mirrors genesis_forge/managers/config/{config_item,params_dict,mdp_fn_class}.py: `fn`, `params`
(a dict that notifies on mutation), and the per-kind attributes `weight` / `time_out` /
`scale` + `noise`.  Values stay LIVE: curricula mutate `cfg[name].weight` or `cfg[name].params[k]`
between steps (docs/guide/managers/reward.md:132-171), and the fused step re-reads them every step
when it packs the term table.
"""
from __future__ import annotations

import inspect
from typing import Callable


class ParamsDict(dict):
    """dict with an on-change callback (params_dict.py:4-22)."""

    def __init__(self, params: dict, on_change: Callable[[], None]):
        super().__init__(params)
        self._on_change = on_change

    def __setitem__(self, key, value):
        super().__setitem__(key, value)
        self._on_change()

    def __delitem__(self, key):
        super().__delitem__(key)
        self._on_change()


class MdpFnClass:
    """Callable term with state; build() runs at environment build and when params change."""

    def __init__(self, env):
        self.env = env

    def build(self):
        pass

    def __call__(self, env, envs_idx):
        pass


class ResetMdpFnClass(MdpFnClass):
    """Class-style EntityManager on_reset item: __call__(env, entity, envs_idx, **params)."""

    def __init__(self, env, entity):
        self.env = env

    def build(self):
        pass

    def __call__(self, env, entity, envs_idx):
        pass


class ConfigItem:
    def __init__(self, cfg: dict, env):
        self._env = env
        self._kwargs = {}
        self._cfg = cfg
        self._fn = cfg["fn"]
        self._params = ParamsDict(cfg.get("params", {}) or {}, self._rebuild)
        self._is_class = inspect.isclass(cfg["fn"])
        self.version = 0  # bumped on every params change; the fused step watches it

    @property
    def fn(self):
        return self._fn

    @property
    def params(self):
        return self._params

    @params.setter
    def params(self, params: dict):
        self._params = ParamsDict(params.copy(), self._rebuild)
        self._rebuild()

    def build(self, **kwargs):
        """Instantiate class-style terms (config_item.py:46-77)."""
        self._kwargs = kwargs
        if self._is_class:
            self._instantiate()

    def execute(self, envs_idx):
        self._fn(self._env, **self._kwargs, envs_idx=envs_idx, **self._params)

    def _instantiate(self):
        params = dict(self._params)
        cls = self._cfg["fn"]
        instance = cls(self._env, **self._kwargs, **params)
        instance.build()
        self._fn = instance

    def _rebuild(self):
        self.version += 1
        if self._is_class and self._kwargs is not None and not inspect.isclass(self._fn):
            self._instantiate()


class TerminationConfigItem(ConfigItem):
    def __init__(self, cfg: dict, env):
        super().__init__(cfg, env)
        self.time_out = cfg.get("time_out", False)


class RewardConfigItem(ConfigItem):
    def __init__(self, cfg: dict, env):
        super().__init__(cfg, env)
        self.weight = cfg.get("weight", 0.0)


class ObservationConfigItem(ConfigItem):
    def __init__(self, cfg: dict, env):
        super().__init__(cfg, env)
        self.scale = cfg.get("scale", 1.0)
        self.noise = cfg.get("noise", None)
