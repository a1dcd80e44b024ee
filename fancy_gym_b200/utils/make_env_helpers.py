"""`make_bb`: step env + wrappers + (phase, basis, MP, tracking law) -> batched black-box env.

Public contract of fancy_gym/utils/make_env_helpers.py:68-136, including what it does to the kwargs
dicts the caller passes in (they are completed in place: `action_dim`, `duration`, `tau`, `learn_tau`,
`tau_bound`, `delay_bound`) and the errors it raises:
  ValueError      sub-trajectory learning together with replanning (:91-92); no RawInterfaceWrapper in the
                  wrapper stack (:63-64)
  AssertionError  `time_limit` disagrees with the MP duration (:176-179)
"""
from __future__ import annotations

from collections.abc import MutableMapping
from typing import Iterable, Optional, Union

import numpy as np

from ..black_box import factory
from ..black_box.black_box_wrapper import BlackBoxWrapper
from ..black_box.raw_interface_wrapper import RawInterfaceWrapper
from . import gym_compat
from .wrappers import TimeAwareObservation


def _stack_has_interface(env) -> bool:
    """walks env -> env.env -> ... looking for a RawInterfaceWrapper"""
    layer = env
    while True:
        if isinstance(layer, RawInterfaceWrapper):
            return True
        if not hasattr(layer, 'env'):
            return False
        layer = layer.env


def _make_wrapped_env(env, wrappers: Iterable[type], seed=1, fallback_max_steps: Optional[int] = None):
    """Applies `wrappers` in order; the result must expose the MP interface somewhere in its stack
    (make_env_helpers.py:35-65).  `fallback_max_steps` gives a step limit to envs registered without one."""
    if fallback_max_steps and not env.spec.max_episode_steps:
        env.spec.max_episode_steps = int(getattr(env.unwrapped, 'max_path_length', fallback_max_steps))
    exposes_interface = _stack_has_interface(env)
    for wrap in wrappers:
        exposes_interface |= issubclass(wrap, RawInterfaceWrapper)
        env = wrap(env)
    if not exposes_interface:
        raise ValueError("A RawInterfaceWrapper is required in order to leverage movement primitive environments.")
    return env


def get_env_duration(env) -> float:
    """episode length in seconds: registered step limit x control period (make_env_helpers.py:148-150)"""
    return env.spec.max_episode_steps * env.dt


def _verify_time_limit(mp_time_limit: Optional[float], env_time_limit: Optional[float]):
    if mp_time_limit is None or env_time_limit is None:
        return
    assert mp_time_limit == env_time_limit, \
        f"The specified 'time_limit' of {env_time_limit}s does not match the duration of {mp_time_limit}s for the MP."


def _complete_phase_kwargs(phase_kwargs: MutableMapping, duration: float, dt: float, sub_trajs):
    """tau defaults to the episode duration; learning sub-trajectories forces a learned tau — also for an explicit
    `learn_sub_trajectories=False`, because the reference tests `is not None` (:115-117);
    a learned tau needs >= 2 env steps (the velocity is a finite difference) and <= one episode, a learned
    delay leaves >= 2 steps (:119-126)."""
    if phase_kwargs.get('tau') is None:
        phase_kwargs['tau'] = duration
    if sub_trajs is not None:
        phase_kwargs['learn_tau'] = True
    if phase_kwargs.get('learn_tau') and phase_kwargs.get('tau_bound') is None:
        phase_kwargs['tau_bound'] = [dt * 2, duration]
    if phase_kwargs.get('learn_delay') and phase_kwargs.get('delay_bound') is None:
        phase_kwargs['delay_bound'] = [0, duration - dt * 2]


def make_bb(env: Union[gym_compat.Env, str], wrappers: Iterable, black_box_kwargs: MutableMapping,
            traj_gen_kwargs: MutableMapping, controller_kwargs: MutableMapping, phase_kwargs: MutableMapping,
            basis_kwargs: MutableMapping, time_limit: int = None, fallback_max_steps: int = None, **kwargs):
    """`env` may be an id (then `**kwargs`, e.g. num_envs= / device=, go to `make`) or a step env."""
    _verify_time_limit(traj_gen_kwargs.get("duration"), time_limit)

    sub_trajs = black_box_kwargs.get('learn_sub_trajectories')
    schedule = black_box_kwargs.get('replanning_schedule')
    if sub_trajs and schedule:
        raise ValueError('Cannot used sub-trajectory learning and replanning together.')

    # a policy that plans more than once per episode has to see the time: the time-aware observation goes in
    # first so that the MP wrapper's context mask still lines up (:94-97; the caller's list is extended, as there)
    if (sub_trajs or schedule) and not any(issubclass(w, TimeAwareObservation) for w in wrappers):
        wrappers.insert(0, TimeAwareObservation)

    step_env = gym_compat.make(env, **kwargs) if isinstance(env, str) else env
    step_env = _make_wrapped_env(step_env, wrappers, fallback_max_steps=fallback_max_steps)

    traj_gen_kwargs.setdefault('action_dim', int(np.prod(step_env.action_space.shape)))
    if black_box_kwargs.get('duration') is None:
        black_box_kwargs['duration'] = get_env_duration(step_env)
    _complete_phase_kwargs(phase_kwargs, black_box_kwargs['duration'], step_env.dt, sub_trajs)

    phase_gen = factory.get_phase_generator(**phase_kwargs)
    basis_gen = factory.get_basis_generator(phase_generator=phase_gen, **basis_kwargs)
    traj_gen = factory.get_trajectory_generator(basis_generator=basis_gen, device=step_env.unwrapped.device,
                                                **traj_gen_kwargs)
    law = factory.get_controller(**controller_kwargs)
    return BlackBoxWrapper(step_env, trajectory_generator=traj_gen, tracking_controller=law, **black_box_kwargs)
