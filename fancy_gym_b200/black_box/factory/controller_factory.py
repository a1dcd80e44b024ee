"""controller_type -> tracking law (fancy_gym/black_box/factory/controller_factory.py:6-21)."""
from ..controller import laws
from ._select import TypeSelector

_SELECT = TypeSelector("controller", {"motor": laws.PDController, "velocity": laws.VelController,
                                      "position": laws.PosController, "metaworld": laws.MetaWorldController})
ALL_TYPES = _SELECT.advertised


def get_controller(controller_type: str, **kwargs):
    return _SELECT.build(controller_type, **kwargs)
