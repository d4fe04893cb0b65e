"""`np_random(seed)` with the semantics of gymnasium >= 0.26: a numpy
Generator over PCG64 seeded through SeedSequence."""
import numpy as np


def np_random(seed=None):
    seed_seq = np.random.SeedSequence(seed)
    np_seed = seed_seq.entropy
    return np.random.Generator(np.random.PCG64(seed_seq)), np_seed
