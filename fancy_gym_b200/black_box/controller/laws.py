"""Tracking laws: (desired pos, desired vel, current pos, current vel) -> action.

Host-side twins of the reference's controller classes (fancy_gym/black_box/controller/
{base,pd,pos,vel,meta_world}_controller.py).  The fused rollout kernel evaluates the same laws per
env from `abi_code` and `gain_vectors()`; these classes exist for the public API (factories, the
reference's controller tests, user code that calls `tracking_controller(...)` directly) and accept
numpy arrays or torch tensors with any leading batch shape.
"""
from __future__ import annotations

import numpy as np

#: fg_ctrl_kind of include/fancy_gym_b200.h for the laws the kernel implements
ABI_CODES = {"velocity": 0, "position": 1, "motor": 2}


def _require_same_shape(what, desired, current):
    ds, cs = tuple(np.shape(desired)), tuple(np.shape(current))
    if ds != cs:
        raise ValueError(f"Mismatch in dimension between desired {what} {ds} and current {what} {cs}")


class BaseController:
    """base_controller.py:4-19 — `get_action` is the law, calling the object is the same thing."""
    kind: str | None = None

    @property
    def abi_code(self):
        """fg_ctrl_kind, or None when the law only exists on the host"""
        return ABI_CODES.get(self.kind)

    def gain_vectors(self, dof: int):
        """(p[dof], d[dof]) float64 as the kernel consumes them; laws without gains report zeros"""
        expand = lambda g: np.broadcast_to(np.asarray(g, dtype=np.float64), (dof,)).copy()  # noqa: E731
        return expand(getattr(self, "p_gains", 0.0)), expand(getattr(self, "d_gains", 0.0))

    def get_action(self, des_pos, des_vel, c_pos, c_vel):
        raise NotImplementedError

    def __call__(self, des_pos, des_vel, c_pos, c_vel):
        return self.get_action(des_pos, des_vel, c_pos, c_vel)


class PDController(BaseController):
    """pd_controller.py:15-29: torque = p (q_des - q) + d (qd_des - qd); gains scalar or per joint;
    ValueError on a desired/current shape mismatch."""
    kind = "motor"

    def __init__(self, p_gains=1, d_gains=0.5):
        self.p_gains, self.d_gains = p_gains, d_gains

    def get_action(self, des_pos, des_vel, c_pos, c_vel):
        _require_same_shape("position", des_pos, c_pos)
        _require_same_shape("velocity", des_vel, c_vel)
        pos_err = des_pos - c_pos
        vel_err = des_vel - c_vel
        return self.p_gains * pos_err + self.d_gains * vel_err


class VelController(BaseController):
    """vel_controller.py:8-9: the desired velocity is the action"""
    kind = "velocity"

    def get_action(self, des_pos, des_vel, c_pos, c_vel):
        return des_vel


class PosController(BaseController):
    """pos_controller.py:8-9: the desired position is the action"""
    kind = "position"

    def get_action(self, des_pos, des_vel, c_pos, c_vel):
        return des_pos


class MetaWorldController(BaseController):
    """meta_world_controller.py:15-25: [xyz delta to the desired position, raw gripper command].
    Host only (Metaworld is outside the fused path)."""
    kind = "metaworld"

    def get_action(self, des_pos, des_vel, c_pos, c_vel):
        target_xyz, gripper = des_pos[:-1], des_pos[-1]
        now_xyz = c_pos[:-1]
        _require_same_shape("position", target_xyz, now_xyz)
        return np.hstack([target_xyz - now_xyz, gripper])
