"""Small host helpers of the solver front end (reference: splitfxm/simhelper.py:7-116)."""
import numpy as np
from scipy.optimize import OptimizeResult


class Solution(OptimizeResult):
    pass


def array_list_reshape(l, shape):
    return np.array(l).reshape(shape)


def report_bounds(y, bounds, t_current):
    """Print the reference's bound warnings (simhelper.py:91-116) for a host copy of the state."""
    lower, upper = np.asarray(bounds[0], dtype=np.float64), np.asarray(bounds[1], dtype=np.float64)
    y = np.asarray(y)
    for idx in np.nonzero(y < lower)[0]:
        print(f"Warning: At t={t_current}, bounds exceeded for variable index {idx}, value = {y[idx]}, "
              f"lower bound = {lower[idx]}")
    for idx in np.nonzero(y > upper)[0]:
        print(f"Warning: At t={t_current}, bounds exceeded for variable index {idx}, value = {y[idx]}, "
              f"upper bound = {upper[idx]}")


apply_bounds = report_bounds
