"""Global seed source, like bayesfast/utils/random.py:8-17 (a numpy Generator) -- it only produces the 64-bit
seeds and x_0 draws; the per-chain streams themselves are the Philox streams of include/bfb_rng.h."""
import numpy as np

__all__ = ['get_generator', 'set_generator', 'new_seed']

_random_generator = np.random.default_rng()


def get_generator():
    return _random_generator


def set_generator(generator):
    global _random_generator
    _random_generator = np.random.default_rng(generator)


def new_seed(generator=None):
    g = get_generator() if generator is None else np.random.default_rng(generator)
    return int(g.integers(0, 2**63 - 1))
