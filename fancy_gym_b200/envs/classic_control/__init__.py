from .mp_wrappers import MPWrapper_HoleReacher, MPWrapper_SimpleReacher, MPWrapper_ViaPointReacher  # noqa: F401
from .reacher import HoleReacherEnv, SimpleReacherEnv, ViaPointReacherEnv  # noqa: F401
