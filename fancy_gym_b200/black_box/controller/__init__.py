from .laws import BaseController, MetaWorldController, PDController, PosController, VelController  # noqa: F401
