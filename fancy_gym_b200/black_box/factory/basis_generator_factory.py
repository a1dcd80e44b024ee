from ...mp import (ExpDecayPhaseGenerator, NormalizedRBFBasisGenerator, PhaseGenerator, ProDMPBasisGenerator,
                   ZeroPaddingNormalizedRBFBasisGenerator)

ALL_TYPES = ["rbf", "zero_rbf", "rhythmic"]


def get_basis_generator(basis_generator_type: str, phase_generator: PhaseGenerator, **kwargs):
    """fancy_gym/black_box/factory/basis_generator_factory.py:8-23"""
    basis_generator_type = basis_generator_type.lower()
    if basis_generator_type == "rbf":
        return NormalizedRBFBasisGenerator(phase_generator, **kwargs)
    elif basis_generator_type == "zero_rbf":
        return ZeroPaddingNormalizedRBFBasisGenerator(phase_generator, **kwargs)
    elif basis_generator_type == "prodmp":
        assert isinstance(phase_generator, ExpDecayPhaseGenerator)
        return ProDMPBasisGenerator(phase_generator, **kwargs)
    elif basis_generator_type == "rhythmic":
        raise NotImplementedError()
    raise ValueError(f"Specified basis generator type {basis_generator_type} not supported, "
                     f"please choose one of {ALL_TYPES}.")
