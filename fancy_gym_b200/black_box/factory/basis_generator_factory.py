"""basis_generator_type -> basis generator (fancy_gym/black_box/factory/basis_generator_factory.py:5-23);
`prodmp` insists on an exponential-decay phase, `rhythmic` is reserved."""
from ... import mp
from ._select import TypeSelector


def _needs_exp_phase(phase_generator, **_):
    assert isinstance(phase_generator, mp.ExpDecayPhaseGenerator)


_SELECT = TypeSelector("basis generator",
                       {"rbf": mp.NormalizedRBFBasisGenerator, "zero_rbf": mp.ZeroPaddingNormalizedRBFBasisGenerator,
                        "prodmp": mp.ProDMPBasisGenerator},
                       reserved=("rhythmic",), advertised=["rbf", "zero_rbf", "rhythmic"],
                       requires={"prodmp": _needs_exp_phase})
ALL_TYPES = _SELECT.advertised


def get_basis_generator(basis_generator_type: str, phase_generator, **kwargs):
    return _SELECT.build(basis_generator_type, phase_generator, **kwargs)
