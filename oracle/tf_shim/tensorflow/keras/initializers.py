import numpy as np


class Constant:
    def __init__(self, value=0.0):
        self.value = value

    def __call__(self, shape, dtype=np.float32):
        return np.full(shape, self.value, dtype=dtype)
