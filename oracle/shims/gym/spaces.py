import numpy as np
class Box:
    def __init__(self, low, high, shape=None, dtype=np.float32):
        low = np.asarray(low, dtype=dtype); high = np.asarray(high, dtype=dtype)
        if shape is not None and low.shape != tuple(shape):
            low = np.full(shape, low, dtype=dtype); high = np.full(shape, high, dtype=dtype)
        self.low, self.high, self.shape, self.dtype = low, high, low.shape, dtype
class Discrete:
    def __init__(self, n): self.n = n
class Tuple:
    def __init__(self, spaces): self.spaces = spaces
class Dict(dict): pass
