"""NumPy / scipy restatement of Simulation.evolve (simulation.py:396-472) over oracle/core.py.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  The reference integrates the flat interior vector with its own
explicit Euler loop (float-accumulated time, a trailing short step, simulation.py:455-463) or hands the right-hand side -
and, for Radau / BDF, the finite-difference Jacobian at y0 as a constant matrix (simulation.py:27, 441-443) - to
scipy.integrate.solve_ivp with t_eval = [t_diff] (simulation.py:466-469).  scipy is installed here, so the restatement
calls the same solve_ivp with the oracle's right-hand side and Jacobian: it pins the oracle's residual / Jacobian along
whole trajectories against the reference's evolve fixtures (tests/test_oracle_golden.py) and is the CPU twin of the device
steppers (splitfxm_b200/ode.py)."""
from __future__ import annotations

import numpy as np

JAC_REQUIRED = ("Radau", "BDF")          # simulation.py:27


def evolve(problem, y0, t_diff, split=False, split_locs=None, method="RK45", max_step=np.inf):
    """-> (y(t_diff) as a flat vector in the run's layout, number of right-hand-side evaluations)."""
    nfev = [0]

    def f(_, y):                                             # simulation.py:435
        nfev[0] += 1
        return problem.residual(y, split, split_locs)

    y = np.array(y0, dtype=np.float64, copy=True)
    if method == "Euler":                                    # simulation.py:455-463
        t_current = 0.0
        while t_current < t_diff:
            delta_t = min(t_diff - t_current, max_step)
            y += delta_t * f(t_current, y)
            t_current += delta_t
        return y, nfev[0]
    from scipy.integrate import solve_ivp
    jac = problem.jacobian_matrix(y, split, split_locs) if method in JAC_REQUIRED else None     # simulation.py:441-443
    sol = solve_ivp(f, (0, t_diff), y, method=method, t_eval=[t_diff], jac=jac, max_step=max_step)   # simulation.py:468-469
    return sol.y.reshape(-1), nfev[0]
