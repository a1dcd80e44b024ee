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
