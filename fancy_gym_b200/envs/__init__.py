"""Registrations of the classic_control envs — fancy_gym/envs/__init__.py:36-87."""
from .classic_control import (HoleReacherEnv, MPWrapper_HoleReacher, MPWrapper_SimpleReacher,
                              MPWrapper_ViaPointReacher, SimpleReacherEnv, ViaPointReacherEnv)
from .registry import (ALL_MOVEMENT_PRIMITIVE_ENVIRONMENTS, MOVEMENT_PRIMITIVE_ENVIRONMENTS_FOR_NS,  # noqa: F401
                       register, upgrade)

register(id='fancy/SimpleReacher-v0', entry_point=SimpleReacherEnv, mp_wrapper=MPWrapper_SimpleReacher,
         max_episode_steps=200, kwargs={"n_links": 2})

register(id='fancy/LongSimpleReacher-v0', entry_point=SimpleReacherEnv, mp_wrapper=MPWrapper_SimpleReacher,
         max_episode_steps=200, kwargs={"n_links": 5})

register(id='fancy/ViaPointReacher-v0', entry_point=ViaPointReacherEnv, mp_wrapper=MPWrapper_ViaPointReacher,
         max_episode_steps=200, kwargs={"n_links": 5, "allow_self_collision": False, "collision_penalty": 1000})

register(id='fancy/HoleReacher-v0', entry_point=HoleReacherEnv, mp_wrapper=MPWrapper_HoleReacher,
         max_episode_steps=200,
         kwargs={"n_links": 5, "random_start": True, "allow_self_collision": False, "allow_wall_collision": False,
                 "hole_width": None, "hole_depth": 1, "hole_x": None, "collision_penalty": 100})


def _register_with_gymnasium():
    """When gymnasium is installed, the same ids are also registered there (`gymnasium.make('fancy_ProMP/HoleReacher-v0',
    num_envs=..., device=...)` then returns the batched black-box env of this package).  gymnasium is not a dependency."""
    try:
        import gymnasium
    except Exception:       # noqa: BLE001
        return False
    from ..utils.gym_compat import registry as own
    for env_id, spec in own.items():
        if env_id in gymnasium.registry:
            continue
        try:
            gymnasium.register(id=env_id, entry_point=spec.entry_point, max_episode_steps=None, kwargs=dict(spec.kwargs),
                               disable_env_checker=True, order_enforce=False)
        except Exception:   # noqa: BLE001
            pass
    return True


_register_with_gymnasium()
