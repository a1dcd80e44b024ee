from .laws import MetaWorldController  # noqa: F401  (import path kept for fancy_gym users)
