"""Nonlinear steady states on the CUDA path against the specification (oracle/newton.py; parity unpinned against the
absent upstream splitnewton): viscous Burgers (custom device model) and the conservative-Euler duct, with active Armijo
back-tracking, bound clipping and pseudo-transient growth (tests/test_newton_nonlinear_oracle.py asserts on CPU that the
specification really takes those branches and >= 4 iterations).  Same iteration counts, iterates within 1e-10 relative:
one mesh, ensembles whose members need different numbers of iterations, and 2-3 slabs."""
import numpy as np
import pytest

from helpers import rel_err
from nonlinear_cases import (BURGERS_BCS, BURGERS_COMPS, BURGERS_NX, BURGERS_SRC, CASES, OracleBurgers, burgers_initial,
                             euler_duct, extend_bounds, oracle_case, oracle_solve, product_case)

pytestmark = pytest.mark.gpu
TOL = 1e-10        # BASELINE.json north_star: Newton iterates within 1e-10 relative


@pytest.mark.parametrize("name", sorted(CASES))
def test_steady_state_single_mesh(name):
    """Simulation.steady_state (simulation.py:474-526) through the Python API -> C ABI."""
    x0, xf, it_ref, trace = oracle_solve(name)
    s, d, c = product_case(name)
    assert np.array_equal(d.listify_interior(c["split"], c["locs"]), x0)
    it = s.steady_state(split=c["split"], split_locs=c["locs"], **c["kw"])
    assert it == it_ref and it >= c["expect"]["min_it"]
    assert rel_err(d.listify_interior(c["split"], c["locs"]), xf) < TOL


def _spec_run(P, U0, split, locs, kw):
    from oracle import core, newton as onewton
    nx, nv = U0.shape
    x0 = core.to_flat(U0, split, locs)
    f = lambda u: P.residual(u, split, locs)
    jac = lambda u: P.jacobian_matrix(u, split, locs)
    okw = dict(kw)
    if "bounds" in okw:
        okw["bounds"] = extend_bounds(okw["bounds"], nx, nv, split, locs)
    if split:
        return onewton.split_newton(f, jac, x0, [nx * i for i in locs], maxiter=200, sparse=True, **okw)
    return onewton.newton(f, jac, x0, maxiter=200, sparse=True, **okw)


def _spec_run_both(P, U0, split, locs, kw):
    """The specification twice - sparse (SuperLU) and dense (LAPACK) linear solves.  The finite-difference Jacobian's
    round-off noise (~1e-6 absolute per entry) is a discontinuous function of the iterate's last bits, so two runs whose
    linear solves round differently drift apart by (noise contraction rate) x (last step) and, on under-resolved shocks
    where the damped iteration wanders, even in their iteration count: nu = 0.03 gives 25 vs 22 iterations between
    these two runs of the SAME specification.  A member is `stable` when the specification agrees with itself."""
    from oracle import core, newton as onewton
    nx, nv = U0.shape
    x0 = core.to_flat(U0, split, locs)
    f = lambda u: P.residual(u, split, locs)
    okw = dict(kw)
    if "bounds" in okw:
        okw["bounds"] = extend_bounds(okw["bounds"], nx, nv, split, locs)
    runs = []
    for dense in (False, True):
        jac = (lambda u: P.jacobian_matrix(u, split, locs).toarray()) if dense else (lambda u: P.jacobian_matrix(u, split, locs))
        if split:
            xf, _, it = onewton.split_newton(f, jac, x0, [nx * i for i in locs], maxiter=200, sparse=not dense, **okw)
        else:
            xf, _, it = onewton.newton(f, jac, x0, maxiter=200, sparse=not dense, **okw)
        runs.append((xf, it))
    (xa, ita), (xb, itb) = runs
    stable = ita == itb and ita < 200 and rel_err(xa, xb) < 1e-12
    return xa, ita, stable


@pytest.mark.parametrize("B", [24, 2048])          # PCR (one thread per cell-member) / block Thomas (one thread per member)
@pytest.mark.parametrize("method,scheme,split,locs,kw", [
    ("FVM", "LAX_FRIEDRICHS", False, None, {}),
    ("FVM", "LAX_FRIEDRICHS", True, [1], {}),
    ("FDM", "CENTRAL", False, None, {"armijo": True}),
])
def test_ensemble_members_take_different_iteration_counts(B, method, scheme, split, locs, kw):
    """A viscosity sweep: the thinner the viscous shock the more Newton iterations - every member must follow the scalar
    specification (its own count, its own Armijo factors) although all members share the kernel launches.  Asserted on
    the members for which the specification is rounding-stable (see _spec_run_both); at least four of the sampled
    members, with at least two different counts, must be."""
    import splitfxm_b200 as sf
    from oracle import core
    from splitfxm_b200.custom import compile_model
    from splitfxm_b200.ensemble import EnsembleSimulation
    nus = np.geomspace(0.036, 0.2, B)
    dm = compile_model(BURGERS_SRC, 2, [0.05], method)
    sch = (sf.FVSchemes if method == "FVM" else sf.FDSchemes)[scheme]
    d = sf.Domain.from_size(BURGERS_NX, 1, 1, BURGERS_COMPS)
    x = np.array(d.positions(interior=True))
    U0 = burgers_initial(x)
    ens = EnsembleSimulation(d, dm, [nus], BURGERS_BCS, sch, method=method)
    x0 = core.to_flat(U0, split, locs)
    Y, its = ens.steady_state(np.repeat(x0[:, None], B, axis=1), split=split, split_locs=locs, maxiter=200, **kw)
    Y, its = Y.cpu().numpy(), its.cpu().numpy()
    assert its.min() >= 3 and its.max() > its.min()                   # genuinely different counts
    checked = []
    for k in sorted(set(int(round(q)) for q in np.linspace(0, B - 1, 8))):
        P = core.Problem(OracleBurgers(nus[k]), method, scheme, x, 1, 1, BURGERS_BCS, BURGERS_COMPS)
        xf, it_ref, stable = _spec_run_both(P, U0, split, locs, kw)
        if stable:
            assert int(its[k]) == it_ref, (k, nus[k])
            assert rel_err(Y[:, k], xf) < TOL, (k, nus[k])
            checked.append(it_ref)
        else:                                                         # same root, whatever path led there
            assert rel_err(Y[:, k], xf) < 1e-6, (k, nus[k])
    assert len(checked) >= 4 and len(set(checked)) >= 2, checked


def test_ensemble_euler_duct_fused_sweep():
    """Euler duct for 2048 members (gamma sweep): the ensemble path with the Jacobian assembled inside the Thomas sweep
    (sfxm_newton_solve_fused) on a genuinely nonlinear problem."""
    import splitfxm_b200 as sf
    from oracle import core
    from splitfxm_b200.ensemble import EnsembleSimulation
    B = 2048
    gammas = np.linspace(1.3, 1.5, B)
    bcs, x, U0 = euler_duct()
    d = sf.Domain.from_size(U0.shape[0], 1, 1, ["rho", "rhou", "E"])
    ens = EnsembleSimulation(d, "euler_cons", {"gamma": gammas}, bcs, sf.FVSchemes.LAX_FRIEDRICHS, method="FVM")
    assert ens.eng.fused_solve_supported(False, None)
    x0 = core.to_flat(U0, False, None)
    Y, its = ens.steady_state(np.repeat(x0[:, None], B, axis=1), maxiter=100)
    Y, its = Y.cpu().numpy(), its.cpu().numpy()
    for k in (0, 700, 2047):
        P = core.Problem(core.Euler1DConservative(gammas[k]), "FVM", "LAX_FRIEDRICHS", x, 1, 1, bcs, ["rho", "rhou", "E"])
        xf, _, it_ref = _spec_run(P, U0, False, None, {})
        assert int(its[k]) == it_ref and it_ref >= 4
        assert rel_err(Y[:, k], xf) < TOL


@pytest.mark.parametrize("G", [2, 3])
@pytest.mark.parametrize("name", ["burgers_fvm_lf_plain", "burgers_fdm_central_armijo", "burgers_fdm_central_split1_bounds",
                                  "burgers_fvm_lf_ptc", "euler_duct_plain"])
def test_slab_newton_nonlinear(name, G):
    """The same cases on G slabs (threads of one process on one GPU): partitioned block solve + all-reduced norms must
    give the specification's iteration count and iterate."""
    import splitfxm_b200 as sf
    from oracle import core
    from splitfxm_b200.custom import compile_model
    from splitfxm_b200.distributed import SlabSimulation
    from splitfxm_b200.spike import ThreadGroup, ThreadTransport
    P, U0, c = oracle_case(name)
    x0, xf, it_ref, _ = oracle_solve(name)
    nx, nv = U0.shape
    split, locs = c["split"], c["locs"]
    sch = (sf.FVSchemes if c["method"] == "FVM" else sf.FDSchemes)[c["scheme"]]
    if c["model"] == "burgers":
        model, params, comps, bcs = compile_model(BURGERS_SRC, 2, [c["nu"]], c["method"]), None, BURGERS_COMPS, BURGERS_BCS
    else:
        model, params, comps = "euler_cons", {"gamma": 1.4}, ["rho", "rhou", "E"]
        bcs = euler_duct()[0]
    x_global = np.linspace(0.0, 1.0, nx)
    kw = dict(c["kw"])

    def body(rank, comm):
        slab = SlabSimulation(x_global, comps, model, params, bcs, sch, c["method"], rank, G)
        slab.transport = ThreadTransport(comm, slab.left, slab.right)
        y = slab.eng.to_device(core.to_flat(U0[slab.lo:slab.hi], split, locs))
        its = slab.steady_state(y, split, locs, comm=comm, maxiter=200, **kw)
        return int(its[0]), core.to_block(y.cpu().numpy()[:, 0], slab.hi - slab.lo, nv, split, locs)

    parts = ThreadGroup(G).run(body)
    assert all(it == it_ref for it, _ in parts)
    got = np.concatenate([u for _, u in parts])
    assert rel_err(got, core.to_block(xf, nx, nv, split, locs)) < TOL
