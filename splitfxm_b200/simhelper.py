"""Small host helpers of the solver front end (reference: splitfxm/simhelper.py:7-116)."""
import numpy as np
from scipy.optimize import OptimizeResult


class Solution(OptimizeResult):
    pass


def array_list_reshape(l, shape):
    return np.array(l).reshape(shape)


def report_bounds(y, bounds, t_current):
    """Print the reference's bound warnings (apply_bounds, simhelper.py:91-116) for a host copy of the state: one line
    per unknown outside its bounds, in index order."""
    lower, upper = np.asarray(bounds[0], dtype=np.float64), np.asarray(bounds[1], dtype=np.float64)
    y = np.asarray(y)
    below = y < lower
    for idx in np.nonzero(below | (y > upper))[0]:
        if below[idx]:
            print(f"Warning: At t={t_current}, bounds exceeded for variable index {idx}, value = {y[idx]}, "
                  f"lower bound = {lower[idx]}")
        else:
            print(f"Warning: At t={t_current}, bounds exceeded for variable index {idx}, value = {y[idx]}, "
                  f"upper bound = {upper[idx]}")


class DeviceBounds:
    """The same report for a state that lives on the device: the comparison runs where the state is, one flag comes
    back per call and only the offending values are copied to the host (an Euler evolve checks every step)."""

    def __init__(self, bounds, device):
        import torch
        self.lower = torch.as_tensor(np.asarray(bounds[0], dtype=np.float64), device=device)
        self.upper = torch.as_tensor(np.asarray(bounds[1], dtype=np.float64), device=device)

    def report(self, y, t_current):
        import torch
        y = y.reshape(-1)
        below = y < self.lower
        bad = below | (y > self.upper)
        if not bool(bad.any()):
            return 0
        idx = torch.nonzero(bad).reshape(-1)
        vals, lo, hi, bl = y[idx].cpu().numpy(), self.lower[idx].cpu().numpy(), self.upper[idx].cpu().numpy(), below[idx].cpu().numpy()
        for k, i in enumerate(idx.cpu().numpy()):
            if bl[k]:
                print(f"Warning: At t={t_current}, bounds exceeded for variable index {i}, value = {vals[k]}, "
                      f"lower bound = {lo[k]}")
            else:
                print(f"Warning: At t={t_current}, bounds exceeded for variable index {i}, value = {vals[k]}, "
                      f"upper bound = {hi[k]}")
        return int(idx.numel())


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
