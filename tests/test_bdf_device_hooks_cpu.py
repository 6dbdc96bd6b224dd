"""Host side of the device BDF (splitfxm_b200/ode.py device_bdf): scipy's BDF forms `self.I - c * self.J` and hands it to
`self.lu`, then calls `self.solve_lu(LU, b)`.  The product replaces I, J, lu and solve_lu so that the constant Jacobian
stays on the device and (I - c J) d = b is solved there as (J - I/c) d = -(b/c).  Here the same hooks are exercised on
CPU with a dense stand-in for the device solve: scipy's own BDF with the explicit matrix must give the same trajectory
and the same number of right-hand-side evaluations (reference call site: simulation.py:441-443, 468-469)."""
import numpy as np
from scipy.integrate import BDF, solve_ivp
from scipy.sparse import csc_matrix, identity


def test_token_algebra_reproduces_scipy_bdf():
    from splitfxm_b200.ode import _DeviceJacobian, _Identity, _ShiftedSystem
    n = 8
    rng = np.random.default_rng(0)
    A = -np.diag(np.arange(1.0, n + 1.0)) + 0.3 * np.diag(rng.random(n - 1), 1) + 0.2 * np.diag(rng.random(n - 1), -1)
    f = lambda t, y: A @ y
    seen = []

    class HookedBDF(BDF):
        def __init__(self, *a, **k):
            super().__init__(*a, **k)
            self.J, self.I, self.LU = _DeviceJacobian(), _Identity(), None
            self.lu = lambda S: S

            def solve_lu(LU, b):
                assert isinstance(LU, _ShiftedSystem)
                seen.append(LU.c)
                return np.linalg.solve(A - np.eye(n) / LU.c, -(b / LU.c))       # the kernel's shifted form
            self.solve_lu = solve_lu

    y0 = np.ones(n)
    a = solve_ivp(f, (0, 2.0), y0, method=HookedBDF, jac=identity(n, format="csc"), t_eval=[2.0])
    b = solve_ivp(f, (0, 2.0), y0, method="BDF", jac=csc_matrix(A), t_eval=[2.0])
    assert a.nfev == b.nfev and a.njev == b.njev
    assert np.max(np.abs(a.y - b.y)) <= 1e-13 * np.max(np.abs(b.y))
    assert len(set(seen)) > 3            # several step sizes / orders: the shift really changes
