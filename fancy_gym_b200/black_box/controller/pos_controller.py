from .laws import PosController  # noqa: F401  (import path kept for fancy_gym users)
