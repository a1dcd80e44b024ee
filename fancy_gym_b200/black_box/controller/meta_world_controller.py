import numpy as np

from .base_controller import BaseController


class MetaWorldController(BaseController):
    """[xyz position delta, raw gripper position] (fancy_gym/black_box/controller/
    meta_world_controller.py:15-25).  Host-side only: Metaworld is outside the fused path."""
    kind = "metaworld"

    def get_action(self, des_pos, des_vel, c_pos, c_vel):
        gripper_pos = des_pos[-1]
        cur_pos = c_pos[:-1]
        xyz_pos = des_pos[:-1]
        if xyz_pos.shape != cur_pos.shape:
            raise ValueError(f"Mismatch in dimension between desired position {xyz_pos.shape} "
                             f"and current position {cur_pos.shape}")
        return np.hstack([(xyz_pos - cur_pos), gripper_pos])
