"""MP wrappers of the three classic_control reachers: `mp_config` and `context_mask` as in
fancy_gym/envs/classic_control/{hole_reacher,viapoint_reacher,simple_reacher}/mp_wrapper.py."""
import numpy as np

from ...black_box.raw_interface_wrapper import RawInterfaceWrapper


class _ReacherMPWrapper(RawInterfaceWrapper):
    @property
    def current_pos(self):
        return self.env.current_pos

    @property
    def current_vel(self):
        return self.env.current_vel


class MPWrapper_HoleReacher(_ReacherMPWrapper):
    mp_config = {
        'ProMP': {
            'controller_kwargs': {'controller_type': 'velocity'},
            'trajectory_generator_kwargs': {'weights_scale': 2},
        },
        'DMP': {
            'controller_kwargs': {'controller_type': 'velocity'},
            'trajectory_generator_kwargs': {'weights_scale': 500},
            'phase_generator_kwargs': {'alpha_phase': 2.5},
        },
        'ProDMP': {},
    }

    @property
    def context_mask(self):
        env = self.env
        return np.hstack([
            [env.random_start] * env.n_links,  # cos
            [env.random_start] * env.n_links,  # sin
            [env.random_start] * env.n_links,  # velocity
            [env.initial_width is None],       # hole width
            [True] * 2,                        # x-y coordinates of target distance
            [False],                           # env steps
        ])


class MPWrapper_ViaPointReacher(_ReacherMPWrapper):
    mp_config = {
        'ProMP': {
            'controller_kwargs': {'controller_type': 'velocity'},
        },
        'DMP': {
            'controller_kwargs': {'controller_type': 'velocity'},
            'trajectory_generator_kwargs': {'weights_scale': 50},
            'phase_generator_kwargs': {'alpha_phase': 2},
        },
        'ProDMP': {},
    }

    @property
    def context_mask(self):
        env = self.env
        return np.hstack([
            [env.random_start] * env.n_links,
            [env.random_start] * env.n_links,
            [env.random_start] * env.n_links,
            [env.initial_via_target is None] * 2,   # x-y coordinates of via point distance
            [True] * 2,                             # x-y coordinates of target distance
            [False],
        ])


class MPWrapper_SimpleReacher(_ReacherMPWrapper):
    mp_config = {
        'ProMP': {
            'controller_kwargs': {'p_gains': 0.6, 'd_gains': 0.075},
        },
        'DMP': {
            'controller_kwargs': {'p_gains': 0.6, 'd_gains': 0.075},
            'trajectory_generator_kwargs': {'weights_scale': 50},
            'phase_generator_kwargs': {'alpha_phase': 2},
        },
        'ProDMP': {},
    }

    @property
    def context_mask(self):
        env = self.env
        return np.hstack([
            [env.random_start] * env.n_links,
            [env.random_start] * env.n_links,
            [env.random_start] * env.n_links,
            [True] * 2,
            [False],
        ])
