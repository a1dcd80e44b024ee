"""make_bb: assembles a (batched) black-box env — fancy_gym/utils/make_env_helpers.py:35-159."""
from __future__ import annotations

from collections.abc import MutableMapping
from typing import Iterable, Type, Union

import numpy as np

from ..black_box.black_box_wrapper import BlackBoxWrapper
from ..black_box.factory.basis_generator_factory import get_basis_generator
from ..black_box.factory.controller_factory import get_controller
from ..black_box.factory.phase_generator_factory import get_phase_generator
from ..black_box.factory.trajectory_generator_factory import get_trajectory_generator
from ..black_box.raw_interface_wrapper import RawInterfaceWrapper
from .gym_compat import Env, Wrapper, make
from .wrappers import TimeAwareObservation


def _make_wrapped_env(env: Env, wrappers: Iterable[Type[Wrapper]], seed=1, fallback_max_steps=None):
    """make_env_helpers.py:35-65: applies the wrappers and insists on a RawInterfaceWrapper."""
    has_black_box_wrapper = False
    head = env
    while hasattr(head, 'env'):
        if isinstance(head, RawInterfaceWrapper):
            has_black_box_wrapper = True
            break
        head = head.env
    for w in wrappers:
        if issubclass(w, RawInterfaceWrapper):
            has_black_box_wrapper = True
        env = w(env)
    if not has_black_box_wrapper:
        raise ValueError("A RawInterfaceWrapper is required in order to leverage movement primitive environments.")
    return env


def make_bb(
        env: Union[Env, str], wrappers: Iterable, black_box_kwargs: MutableMapping, traj_gen_kwargs: MutableMapping,
        controller_kwargs: MutableMapping, phase_kwargs: MutableMapping, basis_kwargs: MutableMapping,
        time_limit: int = None, fallback_max_steps: int = None, **kwargs):
    """Same arguments and side effects on the kwargs dicts as the reference's make_bb
    (make_env_helpers.py:68-136); `**kwargs` go to the step env (num_envs=, device=, env kwargs)."""
    _verify_time_limit(traj_gen_kwargs.get("duration"), time_limit)

    learn_sub_trajs = black_box_kwargs.get('learn_sub_trajectories')
    do_replanning = black_box_kwargs.get('replanning_schedule')
    if learn_sub_trajs and do_replanning:
        raise ValueError('Cannot used sub-trajectory learning and replanning together.')

    wrappers = list(wrappers)
    if (learn_sub_trajs or do_replanning) and not any(issubclass(w, TimeAwareObservation) for w in wrappers):
        wrappers.insert(0, TimeAwareObservation)

    if isinstance(env, str):
        env = make(env, **kwargs)

    env = _make_wrapped_env(env=env, wrappers=wrappers, fallback_max_steps=fallback_max_steps)

    traj_gen_kwargs['action_dim'] = traj_gen_kwargs.get('action_dim', int(np.prod(env.action_space.shape)))

    if black_box_kwargs.get('duration') is None:
        black_box_kwargs['duration'] = get_env_duration(env)
    if phase_kwargs.get('tau') is None:
        phase_kwargs['tau'] = black_box_kwargs['duration']

    if learn_sub_trajs is not None:
        # (sic) also for an explicit False: make_env_helpers.py:115-117
        phase_kwargs['learn_tau'] = True

    if phase_kwargs.get('learn_tau') and phase_kwargs.get('tau_bound') is None:
        phase_kwargs["tau_bound"] = [env.dt * 2, black_box_kwargs['duration']]
    if phase_kwargs.get('learn_delay') and phase_kwargs.get('delay_bound') is None:
        phase_kwargs["delay_bound"] = [0, black_box_kwargs['duration'] - env.dt * 2]

    phase_gen = get_phase_generator(**phase_kwargs)
    basis_gen = get_basis_generator(phase_generator=phase_gen, **basis_kwargs)
    controller = get_controller(**controller_kwargs)
    traj_gen = get_trajectory_generator(basis_generator=basis_gen, device=env.unwrapped.device, **traj_gen_kwargs)

    return BlackBoxWrapper(env, trajectory_generator=traj_gen, tracking_controller=controller, **black_box_kwargs)


def get_env_duration(env: Env):
    return env.spec.max_episode_steps * env.dt       # make_env_helpers.py:148-150


def _verify_time_limit(mp_time_limit, env_time_limit):
    if mp_time_limit is not None and env_time_limit is not None:
        assert mp_time_limit == env_time_limit, \
            f"The specified 'time_limit' of {env_time_limit}s does not match " \
            f"the duration of {mp_time_limit}s for the MP."
