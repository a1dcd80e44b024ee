"""trajectory_generator_type -> MP (fancy_gym/black_box/factory/trajectory_generator_factory.py:4-21);
`prodmp` insists on a ProDMP basis generator."""
from ... import mp
from ._select import TypeSelector


def _needs_prodmp_basis(basis_generator, *_, **__):
    assert isinstance(basis_generator, mp.ProDMPBasisGenerator)


_SELECT = TypeSelector("movement primitive", {"promp": mp.ProMP, "dmp": mp.DMP, "prodmp": mp.ProDMP},
                       advertised=["promp", "dmp", "idmp"], requires={"prodmp": _needs_prodmp_basis})
ALL_TYPES = _SELECT.advertised


def get_trajectory_generator(trajectory_generator_type: str, action_dim: int, basis_generator, **kwargs):
    return _SELECT.build(trajectory_generator_type, basis_generator, action_dim, **kwargs)
