"""Minimal stand-in for the parts of gymnasium that the reference's classic_control
env files, controllers and BlackBoxWrapper import.  TEST INFRASTRUCTURE ONLY: used by
oracle/ref_loader.py to run the *unmodified* reference files from /root/reference in
this container (gymnasium itself is not installed).  Never imported by the product."""
import numpy as np
from . import spaces  # noqa: F401
from . import utils  # noqa: F401
from .utils import seeding


class Env:
    metadata = {}
    spec = None
    render_mode = None
    _np_random = None

    @property
    def np_random(self):
        if self._np_random is None:
            self._np_random, _ = seeding.np_random()
        return self._np_random

    @np_random.setter
    def np_random(self, value):
        self._np_random = value

    @property
    def unwrapped(self):
        return self

    def reset(self, *, seed=None, options=None):
        if seed is not None:
            self._np_random, _ = seeding.np_random(seed)

    def close(self):
        pass


class Wrapper(Env):
    def __init__(self, env):
        self.env = env

    def __getattr__(self, name):
        if name.startswith('_'):
            raise AttributeError(name)
        return getattr(self.env, name)

    @property
    def unwrapped(self):
        return self.env.unwrapped

    def get_wrapper_attr(self, name):
        return getattr(self, name)

    def step(self, action):
        return self.env.step(action)

    def reset(self, **kwargs):
        return self.env.reset(**kwargs)


class ObservationWrapper(Wrapper):
    def reset(self, **kwargs):
        obs, info = self.env.reset(**kwargs)
        return self.observation(obs), info

    def step(self, action):
        obs, rew, term, trunc, info = self.env.step(action)
        return self.observation(obs), rew, term, trunc, info


class TimeLimit(Wrapper):
    """gymnasium.wrappers.TimeLimit semantics: truncated=True on the max_episode_steps-th step."""

    class _Spec:
        def __init__(self, n):
            self.max_episode_steps = n

    def __init__(self, env, max_episode_steps):
        super().__init__(env)
        self._max_episode_steps = max_episode_steps
        self._elapsed_steps = 0
        self.spec = TimeLimit._Spec(max_episode_steps)

    def step(self, action):
        obs, rew, term, trunc, info = self.env.step(action)
        self._elapsed_steps += 1
        if self._elapsed_steps >= self._max_episode_steps:
            trunc = True
        return obs, rew, term, trunc, info

    def reset(self, **kwargs):
        self._elapsed_steps = 0
        return self.env.reset(**kwargs)
