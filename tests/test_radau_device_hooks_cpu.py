"""Host side of the device Radau (splitfxm_b200/ode.py device_radau): scipy's Radau forms `MU_REAL / h * self.I - J` and
`MU_COMPLEX / h * self.I - J`, hands them to `self.lu` and calls `self.solve_lu(LU, b)` with real and complex right-hand
sides.  The product replaces I, J, lu and solve_lu so that the constant Jacobian stays on the device; the complex system is
solved in its real 2s-unknowns-per-cell form by the block-tridiagonal kernels.  Here the hooks and that real form are
exercised on CPU with dense stand-ins (reference call site: simulation.py:441-443, 468-469)."""
import numpy as np
from scipy.integrate import Radau, solve_ivp
from scipy.sparse import csc_matrix, identity


def test_token_algebra_reproduces_scipy_radau():
    from splitfxm_b200.ode import _DeviceJacobian, _Identity, _MuSystem
    n = 8
    rng = np.random.default_rng(0)
    A = -np.diag(np.arange(1.0, n + 1.0)) + 0.3 * np.diag(rng.random(n - 1), 1) + 0.2 * np.diag(rng.random(n - 1), -1)
    f = lambda t, y: A @ y
    seen = {"real": set(), "complex": set()}

    class HookedRadau(Radau):
        def __init__(self, *a, **k):
            super().__init__(*a, **k)
            self.J, self.I = _DeviceJacobian(), _Identity()
            self.LU_real = self.LU_complex = None

            def lu(S):
                self.nlu += 1
                return S
            self.lu = lu

            def solve_lu(LU, b):
                assert isinstance(LU, _MuSystem)
                seen["complex" if isinstance(LU.mu, complex) else "real"].add(LU.mu)
                return np.linalg.solve(LU.mu * np.eye(n) - A, b)
            self.solve_lu = solve_lu

    y0 = np.ones(n)
    a = solve_ivp(f, (0, 2.0), y0, method=HookedRadau, jac=identity(n, format="csc"), t_eval=[2.0])
    b = solve_ivp(f, (0, 2.0), y0, method="Radau", jac=csc_matrix(A), t_eval=[2.0])
    assert a.nfev == b.nfev and a.njev == b.njev and a.nlu == b.nlu
    assert np.max(np.abs(a.y - b.y)) <= 1e-13 * np.max(np.abs(b.y))
    assert len(seen["real"]) > 2 and len(seen["complex"]) > 2       # several step sizes, both kinds of system


def test_real_form_of_the_complex_block_tridiagonal_system():
    """(mu I - J) z = rhs for a cyclic block-tridiagonal J with s x s blocks, solved as the real system the device
    kernel sees: unknowns (Re z, Im z) interleaved per cell, blocks [[B, 0], [0, B]], diagonal [[B, b I], [-b I, B]],
    shift -a, right-hand side -(Re rhs, Im rhs)  (device_radau._ComplexSegment)."""
    rng = np.random.default_rng(1)
    nx, s = 7, 2
    blocks = rng.standard_normal((nx, 3, s, s))                 # slot 0: cell i-1 (wraps), 1: cell i, 2: cell i+1 (wraps)
    blocks[:, 1] -= 4.0 * np.eye(s)
    J = np.zeros((nx * s, nx * s))
    for i in range(nx):
        for d in range(3):
            c = (i + d - 1) % nx
            J[i * s:(i + 1) * s, c * s:(c + 1) * s] += blocks[i, d]
    mu = 1.7 + 2.3j
    rhs = rng.standard_normal(nx * s) + 1j * rng.standard_normal(nx * s)
    want = np.linalg.solve(mu * np.eye(nx * s) - J, rhs)
    J2 = np.zeros((nx, 3, 2 * s, 2 * s))
    J2[:, :, :s, :s] = blocks
    J2[:, :, s:, s:] = blocks
    r = np.arange(s)
    J2[:, 1, r, s + r] += mu.imag
    J2[:, 1, s + r, r] -= mu.imag
    M = np.zeros((nx * 2 * s, nx * 2 * s))
    for i in range(nx):
        for d in range(3):
            c = (i + d - 1) % nx
            M[i * 2 * s:(i + 1) * 2 * s, c * 2 * s:(c + 1) * 2 * s] += J2[i, d]
    rs = rhs.reshape(nx, s)
    f2 = np.concatenate([rs.real, rs.imag], axis=1).reshape(-1)
    v = np.linalg.solve(M - mu.real * np.eye(nx * 2 * s), -f2).reshape(nx, 2 * s)     # (J2 + shift I) v = -f2
    got = (v[:, :s] + 1j * v[:, s:]).reshape(-1)
    assert np.max(np.abs(got - want)) <= 1e-13 * np.max(np.abs(want))
