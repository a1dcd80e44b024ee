from ..controller import MetaWorldController, PDController, PosController, VelController

ALL_TYPES = ["motor", "velocity", "position", "metaworld"]


def get_controller(controller_type: str, **kwargs):
    """fancy_gym/black_box/factory/controller_factory.py:9-21"""
    controller_type = controller_type.lower()
    if controller_type == "motor":
        return PDController(**kwargs)
    elif controller_type == "velocity":
        return VelController(**kwargs)
    elif controller_type == "position":
        return PosController(**kwargs)
    elif controller_type == "metaworld":
        return MetaWorldController(**kwargs)
    raise ValueError(f"Specified controller type {controller_type} not supported, "
                     f"please choose one of {ALL_TYPES}.")
