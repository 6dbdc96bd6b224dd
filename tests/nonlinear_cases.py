"""Nonlinear steady-state problems for the Newton / Split-Newton parity tests (reference call sites
simulation.py:515-523; the solver specification is oracle/newton.py - parity unpinned against upstream splitnewton).

Every case is built twice from the same numbers: as an oracle ``core.Problem`` (CPU specification) and as a
splitfxm_b200 ``Simulation`` (CUDA path).  The cases were chosen so that the specification takes >= 4 iterations and -
where the case says so - back-tracks (Armijo), clips steps at the bounds, or grows a pseudo time step; the oracle-side
test asserts those properties from the specification's trace, so a case can never silently degenerate to the linear
two-iteration path."""
import numpy as np

BURGERS_SRC = """
struct UserModel {                       // u_t + (u^2/2)_x = nu u_xx, every variable independently
  static constexpr int kParams = 1;
  static constexpr bool kDiagonalDFDU = true, kHasDiffusion = true, kDecoupled = false;
  using Scalar = UserModel;
  __device__ static void F(const double* p, const double (&u)[NV], double (&f)[NV]) {
    for (int j = 0; j < NV; ++j) f[j] = 0.5 * (u[j] * u[j]);
  }
  __device__ static void D(const double* p, const double (&u)[NV], double (&d)[NV]) {
    for (int j = 0; j < NV; ++j) d[j] = p[0] * u[j];
  }
  __device__ static double sigma(const double* p, const double (&u)[NV]) {
    double s = 0.0;
    for (int j = 0; j < NV; ++j) s = fmax(s, fabs(u[j]));
    return s;
  }
  __device__ static double eig(const double* p, const double (&u)[NV], int j) { return u[j]; }
};
"""


class OracleBurgers:
    """NumPy twin of BURGERS_SRC for oracle/core.py's residual / Jacobian machinery."""
    name = "burgers"
    diagonal_dfdu = True

    def __init__(self, nu):
        self.nu = nu

    def F(self, U):
        return 0.5 * (U * U)

    def D(self, U):
        return self.nu * U

    def sigma(self, U):
        return np.max(np.abs(U), axis=-1)

    def eig_diag(self, U):
        return U


BURGERS_COMPS = ["a", "b"]
BURGERS_BCS = {"a": {"left": {"dirichlet": 1.0}, "right": {"dirichlet": -1.0}},
               "b": {"left": {"dirichlet": 0.5}, "right": {"neumann": 0.0}}}
BURGERS_NX = 48
BURGERS_BOUNDS = [[-1.35, 0.3], [1.35, 1.0]]          # the initial guess is inside; full Newton steps overshoot


def burgers_initial(x):
    """A viscous shock forming between u = +1 and u = -1 (variable a) and a bump advected out (variable b)."""
    return np.stack([np.cos(np.pi * x) + 0.3 * np.sin(5 * x), 0.5 + 0.4 * np.exp(-200 * (x - 0.5) ** 2)], axis=1)


EULER_COMPS = ["rho", "rhou", "E"]
EULER_NX = 60


def euler_duct(nx=EULER_NX, amp=0.5, u0=2.0, gamma=1.4):
    """Supersonic duct: dirichlet inflow (rho, rho u, E) on the left, neumann-0 outflow, a 50 % bump in the middle.
    Steady state = the uniform inflow state.  -> (bcs, x, U0)"""
    rho0, p0 = 1.0, 1.0
    E0 = p0 / (gamma - 1) + 0.5 * rho0 * u0 * u0
    bcs = {"rho": {"left": {"dirichlet": rho0}, "right": {"neumann": 0.0}},
           "rhou": {"left": {"dirichlet": rho0 * u0}, "right": {"neumann": 0.0}},
           "E": {"left": {"dirichlet": E0}, "right": {"neumann": 0.0}}}
    x = np.linspace(0.0, 1.0, nx)
    b = np.exp(-80 * (x - 0.4) ** 2)
    rho = rho0 * (1 + amp * b)
    u = u0 * (1 - 0.5 * amp * b)
    p = p0 * (1 + amp * b)
    return bcs, x, np.stack([rho, rho * u, p / (gamma - 1) + 0.5 * rho * u * u], axis=1)


# name -> dict(model, method, scheme, nu, split, locs, kw, expect)
# expect: min_it (iterations of the specification), halvings / clips (True: at least one must occur), ptc (dt must grow)
CASES = {
    "burgers_fvm_lf_plain": dict(model="burgers", method="FVM", scheme="LAX_FRIEDRICHS", nu=0.05, split=False, locs=None,
                                 kw={}, expect=dict(min_it=5)),
    "burgers_fvm_lf_split1": dict(model="burgers", method="FVM", scheme="LAX_FRIEDRICHS", nu=0.05, split=True, locs=[1],
                                  kw={}, expect=dict(min_it=6)),
    "burgers_fdm_central_armijo": dict(model="burgers", method="FDM", scheme="CENTRAL", nu=0.05, split=False, locs=None,
                                       kw={"armijo": True}, expect=dict(min_it=10, halvings=True)),
    "burgers_fdm_central_bounds": dict(model="burgers", method="FDM", scheme="CENTRAL", nu=0.05, split=False, locs=None,
                                       kw={"bounds": BURGERS_BOUNDS}, expect=dict(min_it=10, clips=True)),
    "burgers_fdm_central_split1_armijo": dict(model="burgers", method="FDM", scheme="CENTRAL", nu=0.05, split=True, locs=[1],
                                              kw={"armijo": True}, expect=dict(min_it=10, halvings=True)),
    "burgers_fdm_central_split1_bounds": dict(model="burgers", method="FDM", scheme="CENTRAL", nu=0.05, split=True, locs=[1],
                                              kw={"bounds": BURGERS_BOUNDS}, expect=dict(min_it=10, clips=True)),
    "burgers_fvm_lf_ptc": dict(model="burgers", method="FVM", scheme="LAX_FRIEDRICHS", nu=0.02, split=False, locs=None,
                               kw={"dt0": 0.2, "dtmax": 1e6}, expect=dict(min_it=6, ptc=True)),
    "burgers_fvm_lf_all": dict(model="burgers", method="FVM", scheme="LAX_FRIEDRICHS", nu=0.02, split=False, locs=None,
                               kw={"armijo": True, "dt0": 0.2, "dtmax": 1e6, "bounds": BURGERS_BOUNDS},
                               expect=dict(min_it=6, halvings=True, ptc=True)),
    "euler_duct_plain": dict(model="euler_cons", method="FVM", scheme="LAX_FRIEDRICHS", split=False, locs=None, kw={},
                             expect=dict(min_it=4)),
    "euler_duct_armijo": dict(model="euler_cons", method="FVM", scheme="LAX_FRIEDRICHS", split=False, locs=None,
                              kw={"armijo": True}, expect=dict(min_it=4, halvings=True)),
}


def extend_bounds(bounds, nx, nv, split, locs):
    """simulation.py:204-264 for the flat vector layout."""
    from oracle import core
    return [core.to_flat(np.tile(np.asarray(b, dtype=np.float64), (nx, 1)), split, locs) for b in bounds]


def oracle_case(name):
    """-> (P, U0 [nx, nv], case dict)"""
    from oracle import core
    c = CASES[name]
    if c["model"] == "burgers":
        x = np.linspace(0.0, 1.0, BURGERS_NX)
        P = core.Problem(OracleBurgers(c["nu"]), c["method"], c["scheme"], x, 1, 1, BURGERS_BCS, BURGERS_COMPS)
        return P, burgers_initial(x), c
    bcs, x, U0 = euler_duct()
    P = core.Problem(core.Euler1DConservative(1.4), c["method"], c["scheme"], x, 1, 1, bcs, EULER_COMPS)
    return P, U0, c


def oracle_solve(name, maxiter=200, dense=False):
    """The specification's run -> (x0 flat, xf flat, iterations, trace).  dense: LAPACK instead of SuperLU linear solves
    (a different rounding of the same algorithm - used to show that a case is not rounding-sensitive)."""
    from oracle import core, newton as onewton
    P, U0, c = oracle_case(name)
    nx, nv = U0.shape
    split, locs = c["split"], c["locs"]
    x0 = core.to_flat(U0, split, locs)
    f = lambda u: P.residual(u, split, locs)
    jac = (lambda u: P.jacobian_matrix(u, split, locs).toarray()) if dense else (lambda u: P.jacobian_matrix(u, split, locs))
    kw = dict(c["kw"])
    if "bounds" in kw:
        kw["bounds"] = extend_bounds(kw["bounds"], nx, nv, split, locs)
    trace = []
    if split:
        xf, _, it = onewton.split_newton(f, jac, x0, [nx * i for i in locs], maxiter=maxiter, sparse=not dense, trace=trace, **kw)
    else:
        xf, _, it = onewton.newton(f, jac, x0, maxiter=maxiter, sparse=not dense, trace=trace, **kw)
    return x0, xf, it, trace


def product_case(name):
    """-> (Simulation, Domain, case dict) with the Domain holding the case's initial state."""
    import splitfxm_b200 as sf
    from splitfxm_b200.models.euler1d import Euler1DConservative
    c = CASES[name]
    scheme = (sf.FVSchemes if c["method"] == "FVM" else sf.FDSchemes)[c["scheme"]]
    if c["model"] == "burgers":
        d = sf.Domain.from_size(BURGERS_NX, 1, 1, BURGERS_COMPS)
        d.interior_U()[:, :] = burgers_initial(np.array(d.positions(interior=True)))
        s = sf.Simulation(d, burgers_model(c["nu"], c["method"]), {}, BURGERS_BCS, scheme)
        return s, d, c
    bcs, x, U0 = euler_duct()
    d = sf.Domain.from_size(EULER_NX, 1, 1, EULER_COMPS)
    d.interior_U()[:, :] = U0
    s = sf.Simulation(d, Euler1DConservative(1.4, c["method"]), {}, bcs, scheme)
    return s, d, c


def burgers_model(nu, method):
    """The reference's Model protocol (model.py:8; FV/FD equation holders) with the CUDA functor attached."""
    import splitfxm_b200 as sf
    from splitfxm_b200.custom import compile_model
    from splitfxm_b200.equations.fd_transport import FDTransportEquation
    from splitfxm_b200.equations.fv_transport import FVTransportEquation

    class Burgers(sf.Model):
        def __init__(self):
            F = lambda u: 0.5 * np.asarray(u) ** 2
            D = lambda u: nu * np.asarray(u)
            S = lambda u: np.array([0.0])
            self._equation = (FVTransportEquation(F, D, S, lambda u: np.diag(u)) if method == "FVM"
                              else FDTransportEquation(F, D, S))
            self._device = compile_model(BURGERS_SRC, 2, [nu], method)
    return Burgers()
