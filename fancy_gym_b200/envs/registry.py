"""Black-box env ids and the layered MP config — the registry surface of fancy_gym/envs/registry.py.

What callers of the reference rely on (and what is kept here, under the same names):

* `register(id, entry_point, mp_wrapper, ...)` / `upgrade(id, mp_wrapper, ...)` create one black-box id
  `<namespace>_<MP>/<name>` per MP type in ProMP, DMP, ProDMP (registry.py:137-261; an id without a
  namespace lands in `gym_<MP>/`);
* the lookup tables `ALL_MOVEMENT_PRIMITIVE_ENVIRONMENTS[mp_type | 'all']` and
  `MOVEMENT_PRIMITIVE_ENVIRONMENTS_FOR_NS[ns][mp_type | 'all']`;
* `bb_env_constructor` layers   built-in defaults  <  the MP wrapper's `mp_config[mp_type]`  <  the
  register-time override  <  the make-time `mp_config_override`   (registry.py:280-309) with
  `nested_update`, whose quirk — a sub-dict carrying any `*_type` key REPLACES the layer below instead
  of merging into it (:264-277) — is part of the behaviour.

The step env itself is built by this package's own `make` (gymnasium is not available); extra keyword
arguments such as `num_envs=` and `device=` travel through `**kwargs` to it.
"""
from __future__ import annotations

import copy
import importlib
from collections.abc import Mapping, MutableMapping
from typing import Any, Callable, Dict, List, Optional, Union

import numpy as np

from ..black_box.raw_interface_wrapper import RawInterfaceWrapper
from ..utils import gym_compat
from ..utils.make_env_helpers import make_bb

gym_registry = gym_compat.registry


class DefaultMPWrapper(RawInterfaceWrapper):
    """What `upgrade` falls back to (registry.py:21-51): whole observation is the context; position and
    velocity are read from `env.current_pos` / `env.current_vel`."""

    _HINT = ('DefaultMPWrapper was unable to access env.{0}. Please write a custom MPWrapper (recommended) '
             'or expose this attribute directly.')

    @property
    def context_mask(self):
        return np.full(self.env.observation_space.shape, True)

    def _from_env(self, attr):
        assert hasattr(self.env, attr), self._HINT.format(attr)
        return getattr(self.env, attr)

    @property
    def current_pos(self):
        return self._from_env('current_pos')

    @property
    def current_vel(self):
        return self._from_env('current_vel')


# ---- per-MP-type defaults (registry.py:54-125) ---------------------------------------------------
def _layer(traj, phase, basis):
    return {'wrappers': [], 'trajectory_generator_kwargs': traj, 'phase_generator_kwargs': phase,
            'controller_kwargs': {'controller_type': 'motor', 'p_gains': 1.0, 'd_gains': 0.1},
            'basis_generator_kwargs': basis, 'black_box_kwargs': {}}


_BB_DEFAULTS = {
    'ProMP': _layer({'trajectory_generator_type': 'promp'}, {'phase_generator_type': 'linear'},
                    {'basis_generator_type': 'zero_rbf', 'num_basis': 5, 'num_basis_zero_start': 1,
                     'basis_bandwidth_factor': 3.0}),
    'DMP': _layer({'trajectory_generator_type': 'dmp'}, {'phase_generator_type': 'exp'},
                  {'basis_generator_type': 'rbf', 'num_basis': 5}),
    'ProDMP': _layer({'trajectory_generator_type': 'prodmp', 'duration': 2.0, 'weights_scale': 1.0},
                     {'phase_generator_type': 'exp', 'tau': 1.5},
                     {'basis_generator_type': 'prodmp', 'alpha': 10, 'num_basis': 5}),
}

KNOWN_MPS = list(_BB_DEFAULTS)
_LISTS = KNOWN_MPS + ['all']
ALL_MOVEMENT_PRIMITIVE_ENVIRONMENTS: Dict[str, List[str]] = {k: [] for k in _LISTS}
MOVEMENT_PRIMITIVE_ENVIRONMENTS_FOR_NS: Dict[str, Dict[str, List[str]]] = {}


def _resolve(obj_or_path):
    """'module:attr' strings are accepted wherever a class is (registry.py:169-172)"""
    if callable(obj_or_path):
        return obj_or_path
    module, _, attr = obj_or_path.partition(':')
    return getattr(importlib.import_module(module), attr)


def _split_id(env_id: str):
    """'ns/Name-v3' -> ('ns', 'Name-v3'); checks the shape of the id (registry.py:229-241)"""
    pieces = env_id.split('/')
    if len(pieces) > 2:
        raise ValueError('env id can not contain multiple "/".')
    ns, name = (pieces if len(pieces) == 2 else ('gym', pieces[0]))
    stem = name.split('-')
    assert len(stem) >= 2 and stem[-1].startswith('v'), 'Malformed env id, must end in -v{int}.'
    return ns, name


def register_mp(id: str, base_id: str, mp_wrapper, mp_type: str, mp_config_override: Dict[str, Any] = {}):
    """One black-box id for one MP type (registry.py:228-261)."""
    assert mp_type in KNOWN_MPS, 'Unknown mp_type'
    assert id not in ALL_MOVEMENT_PRIMITIVE_ENVIRONMENTS[mp_type], \
        f'The environment {id} is already registered for {mp_type}.'
    ns, name = _split_id(id)
    bb_id = f'{ns}_{mp_type}/{name}'
    gym_compat.register(id=bb_id, entry_point=bb_env_constructor,
                        kwargs=dict(underlying_id=base_id, mp_wrapper=mp_wrapper, mp_type=mp_type,
                                    _mp_config_override_register=mp_config_override))
    per_ns = MOVEMENT_PRIMITIVE_ENVIRONMENTS_FOR_NS.setdefault(ns, {k: [] for k in _LISTS})
    for table in (ALL_MOVEMENT_PRIMITIVE_ENVIRONMENTS, per_ns):
        table[mp_type].append(bb_id)
        table['all'].append(bb_id)


def register_mps(id: str, base_id: str, mp_wrapper, add_mp_types: List[str] = KNOWN_MPS,
                 mp_config_override: Dict[str, Any] = {}):
    for mp_type in add_mp_types:
        register_mp(id, base_id, mp_wrapper, mp_type, mp_config_override.get(mp_type, {}))


def upgrade(id: str, mp_wrapper: RawInterfaceWrapper = DefaultMPWrapper, add_mp_types: List[str] = KNOWN_MPS,
            base_id: Optional[str] = None, mp_config_override: Dict[str, Any] = {}):
    """MP versions for an already registered step env (registry.py:186-220)."""
    register_mps(id, base_id or id, mp_wrapper, add_mp_types, mp_config_override)


def register(id: str, entry_point: Optional[Union[Callable, str]] = None,
             mp_wrapper: RawInterfaceWrapper = DefaultMPWrapper, register_step_based: bool = True,
             add_mp_types: List[str] = KNOWN_MPS, mp_config_override: Dict[str, Any] = {}, **kwargs):
    """Step env + its MP versions (registry.py:137-183); `**kwargs` go to the step registration
    (max_episode_steps=, kwargs=...)."""
    if register_step_based:
        if id in gym_registry:
            print(f'[Info] Gymnasium env with id "{id}" already exists. You should supply register_step_based=False '
                  'or use fancy_gym.upgrade if you only want to register mp versions of an existing env.')
        assert entry_point is not None, 'You need to provide an entry-point, when registering step-based.'
    mp_wrapper = _resolve(mp_wrapper)
    if register_step_based:
        gym_compat.register(id=id, entry_point=entry_point, **kwargs)
    upgrade(id, mp_wrapper, add_mp_types, mp_config_override=mp_config_override)


def nested_update(base: MutableMapping, update):
    """Recursive dict merge, except that an `update` holding a `*_type` key wins wholesale — choosing
    another generator / controller type must not inherit the old type's kwargs (registry.py:264-277)."""
    if any(key.endswith('_type') for key in update):
        return update
    for key, val in update.items():
        base[key] = nested_update(base.get(key, {}), val) if isinstance(val, Mapping) else val
    return base


_SECTIONS = (('traj_gen_kwargs', 'trajectory_generator_kwargs'), ('black_box_kwargs', 'black_box_kwargs'),
             ('controller_kwargs', 'controller_kwargs'), ('phase_kwargs', 'phase_generator_kwargs'),
             ('basis_kwargs', 'basis_generator_kwargs'))


def bb_env_constructor(underlying_id, mp_wrapper, mp_type, mp_config_override={}, _mp_config_override_register={},
                       **kwargs):
    """Entry point of every `<ns>_<MP>/<name>` id (registry.py:280-309)."""
    wrapped = mp_wrapper(gym_compat.make(underlying_id, **kwargs))

    wrapper_cfg = getattr(wrapped, 'mp_config', {}) if hasattr(wrapped, 'mp_config') else {}
    own = copy.deepcopy(wrapper_cfg.get(mp_type, {}))
    inherit = own.pop('inherit_defaults', wrapper_cfg.get('inherit_defaults', True))

    config = copy.deepcopy(_BB_DEFAULTS[mp_type]) if inherit else {}
    for layer in (own, _mp_config_override_register, mp_config_override):
        nested_update(config, copy.deepcopy(layer))

    wrappers = config.pop('wrappers')
    sections = {arg: config.pop(key, {}) for arg, key in _SECTIONS}
    return make_bb(wrapped, wrappers=wrappers, **sections, **config)
