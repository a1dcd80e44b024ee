"""phase_generator_type -> phase generator (fancy_gym/black_box/factory/phase_generator_factory.py:6-23);
`rhythmic` and `smooth` are reserved names there and raise NotImplementedError."""
from ... import mp
from ._select import TypeSelector

_SELECT = TypeSelector("phase generator", {"linear": mp.LinearPhaseGenerator, "exp": mp.ExpDecayPhaseGenerator},
                       reserved=("rhythmic", "smooth"), advertised=["linear", "exp", "rhythmic", "smooth"])
ALL_TYPES = _SELECT.advertised


def get_phase_generator(phase_generator_type, **kwargs):
    return _SELECT.build(phase_generator_type, **kwargs)
