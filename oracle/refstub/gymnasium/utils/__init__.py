from . import seeding  # noqa: F401


class RecordConstructorArgs:
    """gymnasium.utils.RecordConstructorArgs stand-in (only used as a mix-in base)."""

    def __init__(self, **kwargs):
        pass
