from .base_controller import BaseController


class PosController(BaseController):
    """action = desired position (fancy_gym/black_box/controller/pos_controller.py:8-9)."""
    kind = "position"

    def get_action(self, des_pos, des_vel, c_pos, c_vel):
        return des_pos
