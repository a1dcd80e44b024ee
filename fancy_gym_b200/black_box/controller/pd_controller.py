from typing import Tuple, Union

from .base_controller import BaseController


class PDController(BaseController):
    """trq = p_gains * (des_pos - c_pos) + d_gains * (des_vel - c_vel)
    (fancy_gym/black_box/controller/pd_controller.py:21-29)."""
    kind = "motor"

    def __init__(self, p_gains: Union[float, Tuple] = 1, d_gains: Union[float, Tuple] = 0.5):
        self.p_gains = p_gains
        self.d_gains = d_gains

    def get_action(self, des_pos, des_vel, c_pos, c_vel):
        if des_pos.shape != c_pos.shape:
            raise ValueError(f"Mismatch in dimension between desired position {des_pos.shape} "
                             f"and current position {c_pos.shape}")
        if des_vel.shape != c_vel.shape:
            raise ValueError(f"Mismatch in dimension between desired velocity {des_vel.shape} "
                             f"and current velocity {c_vel.shape}")
        return self.p_gains * (des_pos - c_pos) + self.d_gains * (des_vel - c_vel)
