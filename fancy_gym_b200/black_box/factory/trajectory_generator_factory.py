from ...mp import DMP, ProDMP, ProDMPBasisGenerator, ProMP

ALL_TYPES = ["promp", "dmp", "idmp"]


def get_trajectory_generator(trajectory_generator_type: str, action_dim: int, basis_generator, **kwargs):
    """fancy_gym/black_box/factory/trajectory_generator_factory.py:7-21"""
    trajectory_generator_type = trajectory_generator_type.lower()
    if trajectory_generator_type == "promp":
        return ProMP(basis_generator, action_dim, **kwargs)
    elif trajectory_generator_type == "dmp":
        return DMP(basis_generator, action_dim, **kwargs)
    elif trajectory_generator_type == "prodmp":
        assert isinstance(basis_generator, ProDMPBasisGenerator)
        return ProDMP(basis_generator, action_dim, **kwargs)
    raise ValueError(f"Specified movement primitive type {trajectory_generator_type} not supported, "
                     f"please choose one of {ALL_TYPES}.")
