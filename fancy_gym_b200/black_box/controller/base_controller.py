from .laws import BaseController  # noqa: F401  (import path kept for fancy_gym users)
