from ...mp import ExpDecayPhaseGenerator, LinearPhaseGenerator

ALL_TYPES = ["linear", "exp", "rhythmic", "smooth"]


def get_phase_generator(phase_generator_type, **kwargs):
    """fancy_gym/black_box/factory/phase_generator_factory.py:9-23"""
    phase_generator_type = phase_generator_type.lower()
    if phase_generator_type == "linear":
        return LinearPhaseGenerator(**kwargs)
    elif phase_generator_type == "exp":
        return ExpDecayPhaseGenerator(**kwargs)
    elif phase_generator_type in ("rhythmic", "smooth"):
        raise NotImplementedError()
    raise ValueError(f"Specified phase generator type {phase_generator_type} not supported, "
                     f"please choose one of {ALL_TYPES}.")
