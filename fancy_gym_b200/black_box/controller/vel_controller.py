from .laws import VelController  # noqa: F401  (import path kept for fancy_gym users)
