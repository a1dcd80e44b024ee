"""The handful of gymnasium concepts the black-box path relies on (gymnasium is not a dependency:
it is absent from the target image).  Same names and semantics as gymnasium's: `spaces.Box`,
`Env`, `Wrapper` with attribute forwarding / get_wrapper_attr, `EnvSpec`, `register`, `make`.
"""
from __future__ import annotations

import copy
import importlib
from typing import Any, Callable, Dict, Optional, Union

import numpy as np


class Box:
    """gymnasium.spaces.Box (bounds, shape, dtype, sample, contains).  For batched envs the space
    describes ONE env; `sample()` of a batched env's action space returns [num_envs, *shape]."""

    def __init__(self, low, high, shape=None, dtype=np.float32, seed=None, batch: Optional[int] = None):
        self.dtype = np.dtype(dtype)
        low, high = np.asarray(low), np.asarray(high)
        if shape is not None:
            low, high = np.broadcast_to(low, shape), np.broadcast_to(high, shape)
        self.low = low.astype(self.dtype)
        self.high = high.astype(self.dtype)
        self.shape = self.low.shape
        self.batch = batch
        self._rng = np.random.default_rng(seed)

    def seed(self, seed=None):
        self._rng = np.random.default_rng(seed)

    def sample(self):
        """unbounded dims ~ N(0,1), half-bounded ~ bound +- Exp(1), bounded ~ U(low, high) (gymnasium's rule)"""
        shape = self.shape if self.batch is None else (self.batch, *self.shape)
        lo, hi = np.broadcast_to(self.low, shape), np.broadcast_to(self.high, shape)
        lo_f, hi_f = np.isfinite(lo), np.isfinite(hi)
        out = np.empty(shape, dtype=np.float64)
        unb, upp, low_only, both = ~lo_f & ~hi_f, ~lo_f & hi_f, lo_f & ~hi_f, lo_f & hi_f
        out[unb] = self._rng.normal(size=unb.sum())
        out[low_only] = lo[low_only] + self._rng.exponential(size=low_only.sum())
        out[upp] = hi[upp] - self._rng.exponential(size=upp.sum())
        out[both] = self._rng.uniform(lo[both], hi[both])
        return out.astype(self.dtype)

    def contains(self, x) -> bool:
        x = np.asarray(x)
        if x.shape[-len(self.shape):] != self.shape and self.shape != ():
            return False
        return bool(np.all(x >= self.low) and np.all(x <= self.high))

    def __contains__(self, x):
        return self.contains(x)

    def __eq__(self, other):
        return isinstance(other, Box) and self.shape == other.shape and self.dtype == other.dtype \
            and np.array_equal(self.low, other.low) and np.array_equal(self.high, other.high)

    def __repr__(self):
        return f"Box({self.low.min() if self.low.size else ''}, {self.high.max() if self.high.size else ''}, {self.shape}, {self.dtype})"


class spaces:  # namespace alias so that `spaces.Box` reads like the reference
    Box = Box


class EnvSpec:
    def __init__(self, id: str, entry_point=None, max_episode_steps: Optional[int] = None, kwargs: Optional[dict] = None):
        self.id = id
        self.entry_point = entry_point
        self.max_episode_steps = max_episode_steps
        self.kwargs = kwargs or {}


class Env:
    metadata: Dict[str, Any] = {}
    spec: Optional[EnvSpec] = None
    render_mode = None

    @property
    def unwrapped(self):
        return self

    def get_wrapper_attr(self, name):
        return getattr(self, name)

    def close(self):
        pass


class Wrapper(Env):
    def __init__(self, env):
        self.env = env

    def __getattr__(self, name):
        if name.startswith("_") or name == "env":
            raise AttributeError(name)
        return getattr(self.env, name)

    @property
    def spec(self):
        return self.env.spec

    @property
    def unwrapped(self):
        return self.env.unwrapped

    def get_wrapper_attr(self, name):
        return getattr(self, name)

    def reset(self, **kwargs):
        return self.env.reset(**kwargs)

    def step(self, action):
        return self.env.step(action)


registry: Dict[str, EnvSpec] = {}


def register(id: str, entry_point: Union[Callable, str, None] = None, max_episode_steps: Optional[int] = None,
             kwargs: Optional[dict] = None, **_ignored):
    registry[id] = EnvSpec(id, entry_point, max_episode_steps, kwargs or {})


def _load_entry_point(ep):
    if callable(ep):
        return ep
    mod_name, attr = ep.split(":")
    return getattr(importlib.import_module(mod_name), attr)


def make(id: str, **kwargs):
    """gymnasium.make semantics for our registry: spec kwargs (deep-copied) overridden by call kwargs."""
    if id not in registry:
        raise KeyError(f"No registered env with id: {id}")
    spec = registry[id]
    kw = copy.deepcopy(spec.kwargs)
    kw.update(kwargs)
    env = _load_entry_point(spec.entry_point)(**kw)
    try:
        if getattr(env, "spec", None) is None:
            env.spec = spec
    except AttributeError:
        pass
    return env
