class BaseController:
    """fancy_gym/black_box/controller/base_controller.py: get_action(des_pos, des_vel, c_pos, c_vel).
    Works on numpy arrays and torch tensors of any leading batch shape; inside the fused kernel the
    same law is evaluated per env from `kind` and the gains."""
    kind = None

    def get_action(self, des_pos, des_vel, c_pos, c_vel):
        raise NotImplementedError

    def __call__(self, des_pos, des_vel, c_pos, c_vel):
        return self.get_action(des_pos, des_vel, c_pos, c_vel)
