"""The integer restatement of numpy's PCG64 / SeedSequence stream (oracle/np_rng.py: the specification of the device-side
reset sampler) against numpy itself."""
import numpy as np
import pytest

from oracle.np_rng import PCG64, seed_sequence_state

SEEDS = [0, 1, 2, 7, 12345, 2**31 - 1, 2**32 - 1, 2**32, 2**32 + 5, 2**53 + 1, 2**63 - 1, 2**64 - 1]


@pytest.mark.parametrize("seed", SEEDS)
def test_seed_sequence_and_raw_stream(seed):
    assert seed_sequence_state(seed) == [int(x) for x in np.random.SeedSequence(seed).generate_state(4, np.uint64)]
    ref = np.random.PCG64(np.random.SeedSequence(seed))
    mine = PCG64(seed)
    assert [mine.next64() for _ in range(16)] == [int(x) for x in ref.random_raw(16)]


@pytest.mark.parametrize("seed", SEEDS[:8])
def test_distributions_in_the_reference_draw_order(seed):
    g = np.random.Generator(np.random.PCG64(np.random.SeedSequence(seed)))
    m = PCG64(seed)
    for _ in range(20):       # HoleReacher reset: width, direction, x, first joint — repeated on the running stream
        assert g.uniform(0.15, 0.5) == m.uniform(0.15, 0.5)
        assert g.choice([-1, 1]) == [-1, 1][m.choice2()]
        assert g.uniform(0.3 / 2, 3.5) == m.uniform(0.3 / 2, 3.5)
        assert g.uniform(np.pi / 4, 3 * np.pi / 4) == m.uniform(np.pi / 4, 3 * np.pi / 4)
    pair = g.uniform(low=-2.5, high=2.5, size=2)
    assert tuple(pair) == (m.uniform(-2.5, 2.5), m.uniform(-2.5, 2.5))
