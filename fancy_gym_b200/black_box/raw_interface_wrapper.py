from typing import Tuple, Union

import numpy as np

from ..utils.gym_compat import Wrapper


class RawInterfaceWrapper(Wrapper):
    """The contract a step env must satisfy to be driven by movement primitives
    (fancy_gym/black_box/raw_interface_wrapper.py).  Sub-classes carry the class attribute
    `mp_config` (per-MP-type config overrides merged by the registry)."""

    @property
    def context_mask(self) -> np.ndarray:
        """boolean mask over the step observation: which entries form the black-box (context) observation"""
        return np.ones(self.env.observation_space.shape[0], dtype=bool)

    @property
    def current_pos(self) -> Union[float, int, np.ndarray, Tuple]:
        raise NotImplementedError

    @property
    def current_vel(self) -> Union[float, int, np.ndarray, Tuple]:
        raise NotImplementedError

    @property
    def dt(self) -> float:
        return self.env.dt

    def preprocessing_and_validity_callback(self, action, pos_traj, vel_traj, tau_bound: list = None,
                                            delay_bound: list = None):
        return True, pos_traj, vel_traj

    def set_episode_arguments(self, action, pos_traj, vel_traj):
        return pos_traj, vel_traj

    def episode_callback(self, action, pos_traj, vel_traj):
        return True

    def invalid_traj_callback(self, action, pos_traj, vel_traj, return_contextual_obs=None, tau_bound: list = None,
                              delay_bound: list = None):
        return np.zeros(1), 0, True, False, {}
