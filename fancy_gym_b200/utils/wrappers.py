import numpy as np
import torch

from .gym_compat import Box, Wrapper


class TimeAwareObservation(Wrapper):
    """Appends elapsed_steps / max_episode_steps to the observation (fancy_gym/utils/wrappers.py:11-87).
    make_bb inserts it when replanning or learning sub-trajectories.  For the fused envs the extra
    column is produced by the kernel (`time_aware` flag of the handle); this wrapper extends the
    observation space and the reset observation."""
    time_aware = True

    def __init__(self, env, enforce_dtype_float32=False):
        super().__init__(env)
        space = env.observation_space
        if enforce_dtype_float32:
            assert space.dtype == np.float32
        self.observation_space = Box(np.append(space.low, 0.0), np.append(space.high, 1.0), dtype=space.dtype,
                                     batch=space.batch)

    def reset(self, **kwargs):
        obs, info = self.env.reset(**kwargs)
        t = torch.zeros(obs.shape[0], 1, dtype=obs.dtype, device=obs.device)
        return torch.cat([obs, t], dim=1), info
