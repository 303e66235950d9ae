"""
Base wrapper: forwards the environment API to the wrapped environment (or wrapper).

API of genesis_forge/wrappers/wrapper.py:11-121 -- same properties (`dt`, `num_envs`, `scene`, `robot`,
`num_actions`, `action_space`, `num_observations`, `observation_space`, `extras`, `unwrapped`), same
operations (`build`, `step`, `reset`, `get_observations`, `close`), same rule that an environment
which sets `can_be_wrapped = False` refuses further wrapping.  Wrappers sit outside the hot path; they
exist here so that a training script written against the reference keeps its wrapper stack
(SURVEY.md 8(f) rank 4).  The video and skrl wrappers are out of scope (visualisation / another
framework's base class).
"""
from __future__ import annotations

from ..genesis_env import GenesisEnv


class Wrapper:
    env = None
    can_be_wrapped: bool = True

    def __init__(self, env):
        assert env.can_be_wrapped, f"An environment wrapped with {self.__class__.__name__} cannot be wrapped"
        if not isinstance(env, (GenesisEnv, Wrapper)):
            raise ValueError(f"Expected env to be a `GenesisEnv` or `Wrapper` but got {type(env)}")
        self.env = env

    # -- forwarded properties -------------------------------------------------------------------
    dt = property(lambda self: self.env.dt)
    num_envs = property(lambda self: self.env.num_envs)
    scene = property(lambda self: self.env.scene)
    robot = property(lambda self: self.env.robot)
    num_actions = property(lambda self: self.env.num_actions)
    action_space = property(lambda self: self.env.action_space)
    num_observations = property(lambda self: self.env.num_observations)
    observation_space = property(lambda self: self.env.observation_space)
    extras = property(lambda self: self.env.extras)
    unwrapped = property(lambda self: self.env.unwrapped)

    # -- forwarded operations ---------------------------------------------------------------------
    def build(self) -> None:
        self.env.build()

    def step(self, actions):
        return self.env.step(actions)

    def reset(self, env_ids=None):
        return self.env.reset(env_ids)

    def get_observations(self):
        return self.env.get_observations()

    def close(self):
        return self.env.close()
