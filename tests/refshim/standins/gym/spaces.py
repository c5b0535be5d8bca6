import numpy as np


class Box:
    """gym.spaces.Box(low, high, dtype).  `sample()` is fed by the harness (`Box.sampler`) so that the reference's
    hidden reset action (mrp00:411) comes from the same Philox stream the oracle uses.  Samples are float32 VALUES held
    in a float64 array: under NumPy 1.x (the reference's era, gym==0.21) `np.float32 * python_float` is evaluated in
    float64; NumPy 2 (this image) would keep float32.  A float64 carrier reproduces the 1.x arithmetic exactly."""
    sampler = None

    def __init__(self, low, high, shape=None, dtype=np.float32):
        self.low = np.asarray(low, dtype=dtype)
        self.high = np.asarray(high, dtype=dtype)
        self.shape = self.low.shape if shape is None else tuple(shape)
        self.dtype = np.dtype(dtype)

    def sample(self):
        if Box.sampler is None:
            raise RuntimeError("refshim: Box.sampler not installed by the harness")
        a = np.asarray(Box.sampler(self.shape), dtype=np.float32)
        return a.astype(np.float64)
