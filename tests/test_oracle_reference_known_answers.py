"""The known answers the reference's OWN tests hold for the hot path, evaluated with the CPU oracle:
tests/test_flux.py:54-155 (Lax-Friedrichs / upwind / central face fluxes, diffusion fluxes with the 0.5 face spacing),
tests/test_derivative.py:17-84 (first / second derivatives, CENTRAL and RIGHT_BIAS) and tests/test_bc.py:11-75 (periodic,
outflow, dirichlet, neumann ghost cells).  Same inputs, same expected values as the reference's assertions."""
import numpy as np
import pytest

from oracle import core


class _Linear:
    """F = f u, D = d u with a diagonal dFdU = f I - the flux / diffusion functions of the reference's flux tests."""
    diagonal_dfdu = True

    def __init__(self, f, d=0.5):
        self.f, self.d = f, d

    def F(self, U):
        return U * self.f

    def D(self, U):
        return U * self.d

    def sigma(self, U):
        return np.full(U.shape[0], abs(self.f))

    def eig_diag(self, U):
        return np.full(U.shape, float(self.f))


def _stencil(vals):
    return [np.array([v], dtype=np.float64) for v in vals]


CELLS3 = _stencil([[1.0, 2.0], [3.0, 4.0], [2.0, 6.0]])                   # tests/test_flux.py:33-41
CELLS3_N = _stencil([[1.0, 2.0], [-3.0, 4.0], [2.0, -6.0]])               # tests/test_flux.py:44-51
X3 = [np.array([0.0]), np.array([1.0]), np.array([2.0])]


def test_fluxes_lax_friedrichs():                                         # tests/test_flux.py:54-65
    m = _Linear(2.0)
    ul, uc, ur = CELLS3
    Fw, Fe = core.fv_fluxes(m, "LAX_FRIEDRICHS", ul, uc, ur)
    assert np.allclose(Fw, 0.5 * (m.F(ul) + m.F(uc)) - 0.5 * 2 * (uc - ul))
    assert np.allclose(Fe, 0.5 * (m.F(uc) + m.F(ur)) - 0.5 * 2 * (ur - uc))


def test_fluxes_upwind_both_signs():                                      # tests/test_flux.py:74-107
    ul, uc, ur = CELLS3
    m = _Linear(2.0)
    Fw, Fe = core.fv_fluxes(m, "UPWIND", ul, uc, ur)
    assert np.allclose(Fw, m.F(ul)) and np.allclose(Fe, m.F(uc))
    ul, uc, ur = CELLS3_N
    m = _Linear(-2.0)
    Fw, Fe = core.fv_fluxes(m, "UPWIND", ul, uc, ur)
    assert np.allclose(Fw, m.F(uc)) and np.allclose(Fe, m.F(ur))


def test_fluxes_central_and_unsupported_scheme():                         # tests/test_flux.py:116-140
    m = _Linear(2.0)
    ul, uc, ur = CELLS3
    Fw, Fe = core.fv_fluxes(m, "CENTRAL", ul, uc, ur)
    assert np.allclose(Fw, 0.5 * (m.F(ul) + m.F(uc))) and np.allclose(Fe, 0.5 * (m.F(uc) + m.F(ur)))
    with pytest.raises(core.OracleError):
        core.fv_fluxes(m, "UNSUPPORTED_SCHEME", ul, uc, ur)


def test_diffusion_fluxes_use_half_the_cell_spacing():                    # tests/test_flux.py:143-158
    m = _Linear(2.0, 0.5)
    ul, uc, ur = CELLS3
    Dw, De = core.fv_diffusion_fluxes(m, X3[0], X3[1], X3[2], ul, uc, ur)
    dxw, dxe = 0.5 * (1.0 - 0.0), 0.5 * (2.0 - 1.0)
    assert np.allclose(Dw, (m.D(uc) - m.D(ul)) / dxw) and np.allclose(De, (m.D(ur) - m.D(uc)) / dxe)


@pytest.mark.parametrize("scheme,n", [("CENTRAL", 3), ("RIGHT_BIAS", 4)])
def test_derivatives_of_a_linear_profile(scheme, n):                      # tests/test_derivative.py:17-84
    xs = [np.array([float(k)]) for k in range(n)]
    vals = [np.array([[float(k + 1)]]) for k in range(n)]                 # values 1, 2, 3(, 4) at x = 0, 1, 2(, 3)
    assert np.allclose(core.fd_dx(scheme, xs, vals), [[1.0]])
    assert core.fd_d2x(scheme, xs, vals)[0, 0] == 0.0


def test_stencil_size_must_match():                                       # tests/test_derivative.py:87-91, test_flux.py:128-132
    x_all = np.array([0.0, 1.0])
    U_all = np.array([[1.0], [2.0]])
    with pytest.raises(core.OracleError):
        core.residual_block(_Linear(1.0), "FDM", "CENTRAL", x_all, U_all, 1, 0)


def _gaussian_domain():
    """Domain.from_size(10, 1, 1, ["u", "v", "w"]) with u = gaussian (tests/test_bc.py:13-16)."""
    x = np.linspace(0.0, 1.0, 10)
    U = np.zeros((10, 3))
    U[:, 0] = core.initial_condition(x, "gaussian")
    return x, U


def _ghosts(bcs_u):
    x, U = _gaussian_domain()
    bc_list = core.parse_bcs({"u": bcs_u}, ["u", "v", "w"])
    return core.fill_ghosts(x, U, 1, 1, bc_list, 0.0, 1.0), (x, U)


def test_boundary_conditions():                                           # tests/test_bc.py:11-75
    (xa, Ua), (x, U) = _ghosts({"left": "periodic", "right": "periodic"})
    assert Ua[0, 0] == U[-1, 0] and Ua[-1, 0] == U[0, 0]
    (xa, Ua), (x, U) = _ghosts({"left": "outflow", "right": "outflow"})
    dy, dx = U[1, 0] - U[0, 0], x[1] - x[0]
    assert Ua[0, 0] == U[0, 0] + (dy / dx) * (xa[0] - x[0])
    dy, dx = U[-1, 0] - U[-2, 0], x[-1] - x[-2]
    assert Ua[-1, 0] == U[-1, 0] + (dy / dx) * (xa[-1] - x[-1])
    (xa, Ua), (x, U) = _ghosts({"left": {"dirichlet": 2.0}, "right": {"neumann": 1.0}})
    assert Ua[0, 0] == 2.0
    assert (Ua[-1, 0] - U[-1, 0]) / (xa[-1] - x[-1]) == pytest.approx(1.0)
    (xa, Ua), (x, U) = _ghosts({"left": {"neumann": 2.0}, "right": {"dirichlet": 1.0}})
    assert (U[0, 0] - Ua[0, 0]) / (x[0] - xa[0]) == pytest.approx(2.0)
    assert Ua[-1, 0] == 1.0
    with pytest.raises(core.OracleError):
        _ghosts({"left": "unsupported", "right": "periodic"})


def test_readme_driver_probe_values():
    """SURVEY.md 8(c) probe goldens of the README driver (FVM, nx = 20, u/v/w, split [1, 2]; main.py:14-29 of the reference's
    README): |r|_2, three Jacobian entries, and the state sum and RHS count after evolve(0.1)."""
    from oracle import evolve as oevolve
    comps = ["u", "v", "w"]
    bcs = {"u": {"left": "periodic", "right": "periodic"},
           "v": {"left": {"dirichlet": 3}, "right": {"dirichlet": 4}},
           "w": {"left": {"dirichlet": 2}, "right": "periodic"}}
    x = np.linspace(0.0, 1.0, 20)
    U = np.stack([core.initial_condition(x, k) for k in ("gaussian", "rarefaction", "tophat")], axis=1)
    P = core.Problem(core.AdvectionDiffusion(0.2, 0.001), "FVM", "LAX_FRIEDRICHS", x, 1, 1, bcs, comps)
    x0 = core.to_flat(U, True, [1, 2])
    r = P.residual(x0, True, [1, 2])
    assert float(np.linalg.norm(r)) == 15.60848741616531
    J = P.jacobian_matrix(x0, True, [1, 2]).toarray()
    assert [J[0, 0], J[0, 1], J[0, 19]] == [-5.244000000009059, 0.7219999999990081, 4.521999999999209]
    y, nfev = oevolve.evolve(P, core.to_flat(U, False, None), 0.1)
    xa, Ua = P.ghosts(core.to_block(y, 20, 3, False, None))
    assert nfev == 32
    assert abs(float(Ua.sum()) - 49.6025148316751) <= 1e-12
