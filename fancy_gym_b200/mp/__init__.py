from .basis_gn import (NormalizedRBFBasisGenerator, ProDMPBasisGenerator,  # noqa: F401
                       ZeroPaddingNormalizedRBFBasisGenerator)
from .mp import DMP, MPInterface, ProDMP, ProMP  # noqa: F401
from .phase_gn import ExpDecayPhaseGenerator, LinearPhaseGenerator, PhaseGenerator  # noqa: F401
