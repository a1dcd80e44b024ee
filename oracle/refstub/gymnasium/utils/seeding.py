import numpy as np


def np_random(seed=None):
    """Same construction as gymnasium.utils.seeding.np_random: PCG64 seeded by SeedSequence(seed)."""
    seed_seq = np.random.SeedSequence(seed)
    return np.random.Generator(np.random.PCG64(seed_seq)), seed_seq.entropy
