"""Batched twin of the ToyEnv / ToyWrapper the reference's tests define (test/test_black_box.py:27-56,
test/test_replanning_sequencing.py): obs -1, reward 1, never terminates, dt 0.02; current_pos 1, current_vel 0.
The dynamics run inside the fused kernel as FG_ENV_TOY."""
import numpy as np
import torch

from fancy_gym_b200 import _lib
from fancy_gym_b200.black_box.raw_interface_wrapper import RawInterfaceWrapper
from fancy_gym_b200.utils.gym_compat import Box, Env


class ToyEnv(Env):
    env_kind = _lib.ENV_TOY
    dt = 0.02

    def __init__(self, a: int = 0, b: float = 0.0, c: list = [], d: dict = {}, e=None, num_envs: int = 1, device=None,
                 n_links: int = 1):
        self.a, self.b, self.c, self.d, self.e = a, b, c, d, e
        self.num_envs, self.n_links = int(num_envs), int(n_links)
        self.device = torch.device(device) if device is not None else torch.device("cuda", 0)
        batch = self.num_envs if self.num_envs > 1 else None
        self.observation_space = Box(low=-1, high=1, shape=(1,), dtype=np.float64, batch=batch)
        self.action_space = Box(low=-1, high=1, shape=(self.n_links,), dtype=np.float64, batch=batch)
        B, n, dev = self.num_envs, self.n_links, self.device
        self.q = torch.ones(B, n, dtype=torch.float64, device=dev)
        self.v = torch.zeros(B, n, dtype=torch.float64, device=dev)
        self.steps = torch.zeros(B, dtype=torch.int32, device=dev)
        self.done = torch.zeros(B, dtype=torch.uint8, device=dev)
        self.ctx = torch.zeros(B, 4, dtype=torch.float64, device=dev)

    def reset(self, *, seed=None, options=None):
        self.steps.zero_(); self.done.zero_()
        return -torch.ones(self.num_envs, 1, dtype=torch.float32, device=self.device), {}

    def render(self):
        pass


class ToyWrapper(RawInterfaceWrapper):
    @property
    def current_pos(self):
        return np.ones(self.action_space.shape)

    @property
    def current_vel(self):
        return np.zeros(self.action_space.shape)


def register_toy(fancy_gym, max_episode_steps=50):
    from fancy_gym_b200.utils import gym_compat
    if "toy-v0" not in gym_compat.registry:
        gym_compat.register(id="toy-v0", entry_point=ToyEnv, max_episode_steps=max_episode_steps)
