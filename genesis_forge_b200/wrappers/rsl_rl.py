"""
rsl_rl adapter (genesis_forge/wrappers/rsl_rl.py:10-119): `step` returns (obs, rewards, dones, extras)
with dones = terminated | truncated, the observations are added to `extras["observations"]["critic"]`
when no critic group exists, and with rsl_rl >= 3 observations travel as a TensorDict of groups.

Glue on top of the fused step: `dones` is not computed here -- the post-physics kernel writes
terminated | truncated next to the two masks (GFB_B_DONES), so the wrapper adds no launch to a step;
`extras["time_outs"]` is already published by the environment (termination_manager.py:188-189).
"""
from __future__ import annotations

from importlib import metadata

from .._gs import gs
from ..fused import make_obs_dict
from .wrapper import Wrapper


class RslRlWrapper(Wrapper):
    can_be_wrapped = False

    def __init__(self, env):
        super().__init__(env)
        self.rsl3 = False
        try:
            self.rsl3 = int(metadata.version("rsl-rl-lib").split(".")[0]) >= 3
        except Exception:
            pass

    @property
    def device(self):
        return gs.device

    def step(self, actions):
        obs, rewards, terminated, truncated, extras = super().step(actions)
        base = self.unwrapped
        # the kernel's own OR of the two masks; environments without the fused step fall back to torch
        dones = base.dones if getattr(base, "_fused", None) is not None else terminated | truncated
        extras = self._add_observations_to_extras(obs, extras if extras is not None else {})
        return self._format_obs_group(obs, extras), rewards, dones, extras

    def reset(self):
        obs, extras = self.env.reset()
        return self._format_obs_group(obs, extras), extras

    def get_observations(self):
        obs = self.env.get_observations()
        if self.rsl3:
            return self._format_obs_group(obs, self.env.extras)
        return obs, self._add_observations_to_extras(obs, self.env.extras)

    def _add_observations_to_extras(self, obs, extras):
        if extras is None:
            extras = {}
        if "observations" not in extras:
            extras["observations"] = {}
        if "critic" not in extras["observations"]:
            extras["observations"]["critic"] = obs
        return extras

    def _format_obs_group(self, obs, extras):
        if not self.rsl3:
            return obs
        groups = make_obs_dict(gs.device)
        if extras is not None and "observations" in extras:
            for name, value in extras["observations"].items():
                groups[name] = value
        else:
            groups["policy"] = obs
        return groups
