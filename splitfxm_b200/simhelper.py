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


def create_bound_events(y0, bounds):
    """One non-terminal scipy event per unknown that prints the reference's warning when the unknown leaves its
    bounds (simhelper.py:31-88); [] when there are no bounds."""
    if bounds is None:
        return []
    lower, upper = bounds[0], bounds[1]

    def make(index, lo, hi):
        def event(t, y):
            if y[index] < lo:
                print(f"Warning: At t={t}, bounds exceeded for variable index {index}, value = {y[index]}, lower bound = {lo}")
                return y[index] - lo
            if y[index] > hi:
                print(f"Warning: At t={t}, bounds exceeded for variable index {index}, value = {y[index]}, upper bound = {hi}")
                return hi - y[index]
            return 1
        event.terminal = False
        event.direction = 0
        return event
    return [make(i, lower[i], upper[i]) for i in range(len(y0))]
