import numpy as np


def np_random(seed=None):
    """gym.utils.seeding.np_random: the env's own RNG.  The reference never draws spawns from it (it uses the global
    np.random), so its stream does not influence any compared quantity."""
    return np.random.RandomState(seed if seed is not None else 0), seed
