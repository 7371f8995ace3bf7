import numpy as np

from . import _t


def normal(shape, dtype=np.float32, seed=None):
    return _t(np.random.default_rng(seed).normal(size=tuple(shape)).astype(dtype))
