import numpy as np


class Box:
    def __init__(self, low, high, shape=None, dtype=np.float32):
        self.dtype = np.dtype(dtype)
        low = np.asarray(low)
        high = np.asarray(high)
        if shape is not None:
            low = np.broadcast_to(low, shape)
            high = np.broadcast_to(high, shape)
        self.low = low.astype(self.dtype)
        self.high = high.astype(self.dtype)
        self.shape = self.low.shape


class Dict(dict):
    """placeholder: the classic_control envs never use Dict observations"""
    @property
    def spaces(self):
        return self


def flatten(space, x):
    return x


def flatten_space(space):
    return space
