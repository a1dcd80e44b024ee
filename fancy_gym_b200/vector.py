"""gymnasium.vector.VectorEnv-shaped adapter over the batched black-box env.

RL libraries drive vector envs through  reset(seed=, options=) -> (obs[N, O], infos)  and
step(actions[N, P]) -> (obs, rewards, terminations, truncations, infos)  on numpy arrays, with finished
sub-envs reset automatically (gymnasium <= 0.29 semantics: the returned observation of a finished sub-env is the first
one of its NEXT episode and the last one of the finished episode goes to infos["final_observation"]).  A black-box
step is a whole episode, so every sub-env finishes on every step and the auto-reset is one extra fg_reset launch.

gymnasium itself is not a dependency; the attribute names follow its VectorEnv (num_envs, single_action_space,
single_observation_space, action_space, observation_space, closed).
"""
from __future__ import annotations

from typing import Any, Dict, Optional

import numpy as np
import torch

from .utils.gym_compat import Box, make


class BlackBoxVectorEnv:
    def __init__(self, env_id: str, num_envs: int, device="cuda:0", tensors: bool = False, **make_kwargs):
        """tensors=True keeps everything on the device (torch tensors in, torch tensors out: no host round trip)."""
        self.env = make(env_id, num_envs=num_envs, device=device, **make_kwargs)
        if self.env.do_replanning or self.env.learn_sub_trajectories:
            raise NotImplementedError("BlackBoxVectorEnv auto-resets after every black-box step; replanning / "
                                      "sub-trajectory envs keep one clock per batch and are driven through reset()/step()")
        self.num_envs = int(num_envs)
        self.tensors = bool(tensors)
        self.single_action_space = Box(self.env.action_space.low, self.env.action_space.high, dtype=self.env.action_space.dtype)
        self.single_observation_space = Box(self.env.observation_space.low, self.env.observation_space.high,
                                            dtype=self.env.observation_space.dtype)
        self.action_space = Box(self.single_action_space.low, self.single_action_space.high,
                                dtype=self.single_action_space.dtype, batch=self.num_envs)
        self.observation_space = Box(self.single_observation_space.low, self.single_observation_space.high,
                                     dtype=self.single_observation_space.dtype, batch=self.num_envs)
        self.closed = False
        self.spec = self.env.spec

    def _out(self, x):
        return x if self.tensors or not torch.is_tensor(x) else x.cpu().numpy()

    def reset(self, *, seed: Optional[int] = None, options: Optional[Dict[str, Any]] = None):
        opts = dict(options or {})
        opts["as_numpy"] = False
        obs, info = self.env.reset(seed=seed, options=opts)
        return self._out(obs if self.num_envs > 1 or obs.dim() == 2 else obs[None]), info

    def step(self, actions):
        a = actions if torch.is_tensor(actions) else torch.as_tensor(np.asarray(actions))
        if a.dim() == 1:
            a = a[None]
        obs, ret, terminated, truncated, infos = self.env.step(a.to(self.env.device))
        done = terminated | truncated
        infos = dict(infos)
        infos["final_observation"] = obs
        infos["_final_observation"] = done
        next_obs = self.env.reset_done()           # every finished sub-env starts its next episode (its own context stream)
        return (self._out(next_obs), self._out(ret), self._out(terminated), self._out(truncated),
                {k: self._out(v) for k, v in infos.items()})

    def close(self):
        if not self.closed:
            self.env.close()
            self.closed = True


def make_vec(env_id: str, num_envs: int, **kwargs) -> BlackBoxVectorEnv:
    return BlackBoxVectorEnv(env_id, num_envs, **kwargs)
