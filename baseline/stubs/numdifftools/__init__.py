"""Stand-in for `numdifftools` (absent from this image, no network): the reference imports it at
simulation.py:2 but only its tests/benchmark call it (forward-difference dense Jacobian, step 1e-8)."""
import numpy as np


class Jacobian:
    def __init__(self, f, method="forward", step=1e-8):
        self.f, self.h = f, step

    def __call__(self, x):
        x = np.asarray(x, dtype=float)
        f0 = np.asarray(self.f(x))
        J = np.zeros((f0.size, x.size))
        for k in range(x.size):
            xp = x.copy()
            xp[k] += self.h
            J[:, k] = (np.asarray(self.f(xp)) - f0) / self.h
        return J
