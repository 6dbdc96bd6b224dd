"""Initial conditions (reference: splitfxm/initialize.py:7-59).

Evaluated once, on the host, cell by cell with NumPy scalar arithmetic exactly as the reference does, so the
starting state is bit-identical (np.sin / np.exp on scalars and on arrays may differ in the last bit)."""
import numpy as np

from .domain import Domain


def set_initial_condition(d: Domain, v: str, type="gaussian"):
    idx = d.component_index(v)
    xmin = d.xmin()
    xmax = d.xmax()
    width = xmax - xmin
    xs = d.interior_x()
    col = d.interior_U()[:, idx]
    if type == "tophat":
        for k in range(xs.shape[0]):
            x = xs[k]
            col[k] = 1.0 if (x >= xmin + (0.333 * width) and x <= xmin + (0.666 * width)) else 0.0
    elif type == "sine":
        for k in range(xs.shape[0]):
            x = xs[k]
            col[k] = (1.0 + 0.5 * np.sin(2.0 * np.pi * (x - xmin - 0.333 * width) / (0.333 * width))
                      if (x >= xmin + (0.333 * width) and x <= xmin + (0.666 * width)) else 1.0)
    elif type == "rarefaction":
        for k in range(xs.shape[0]):
            col[k] = 2.0 if xs[k] > xmin + (0.5 * width) else 1.0
    elif type == "gaussian":
        for k in range(xs.shape[0]):
            col[k] = np.exp(-200 * (xs[k] - xmin - 0.25 * width) ** 2.0)
    else:
        raise NotImplementedError
