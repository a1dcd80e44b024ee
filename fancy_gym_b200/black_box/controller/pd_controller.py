from .laws import PDController  # noqa: F401  (import path kept for fancy_gym users)
