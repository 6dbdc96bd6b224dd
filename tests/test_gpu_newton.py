"""Linear solves and Newton / Split-Newton on the device against the oracle specification (oracle/newton.py;
parity unpinned against upstream splitnewton - see that file's header)."""
import numpy as np
import pytest

from helpers import load, oracle_problem, product_simulation, rel_err

pytestmark = pytest.mark.gpu


def _dense_from_banded(eng, Jh, split, locs):
    rows, cols, vals = eng.jacobian_coo(Jh, split, locs)
    n = eng.n
    A = np.zeros((n, n))
    A[rows, cols] = vals
    return A


@pytest.mark.parametrize("name,split,locs", [
    ("ad_fvm_lf_nx40_mixedbc", False, None),          # w periodic on the left only -> cyclic 3x3 blocks
    ("ad_fvm_lf_nx40_mixedbc", True, [1, 2]),         # three scalar systems, one cyclic
    ("ad_fdm_central_nonuniform_mixed", True, [2]),   # 2x2 + 1x1 segments on an AMR mesh
    ("sod_cons_nx60_unsplit_t0p02", False, None),     # dense 3x3 blocks (Euler)
    ("sod_cons_nx100_t0p05", True, [1]),
    ("ad_fvm_lf_nonuniform", True, [1, 2]),           # README BCs: u fully cyclic, v plain, w right-periodic only
    ("ad_fvm_lf_nonuniform", True, [1]),              # (v, w) 2x2 blocks with a lower-left corner block only
    ("ad_fdm_central_nx250_split1", True, [1]),
])
@pytest.mark.parametrize("method", ["thomas", "pcr"])
def test_block_tridiag_solve(name, split, locs, method):
    import torch
    z = load(name)
    s, d, m = product_simulation(z)
    eng = s._engine()
    from oracle import core
    y_host = core.to_flat(core.to_block(z["x0"], m["nx"], len(m["comps"]), m["split"], m["locs"]), split, locs)
    y = eng.to_device(y_host)
    r = torch.empty_like(y)
    J = eng.jacobian(y, split, locs, r0=r)
    shift = torch.full((1,), -7.5, dtype=torch.float64, device=y.device)
    sol = torch.zeros_like(y)
    status = torch.zeros(1, dtype=torch.int32, device=y.device)
    eng.solve(J, r, sol, split, locs, shift=shift, status=status, method=method)
    A = _dense_from_banded(eng, J.cpu().numpy()[..., 0], split, locs) - 7.5 * np.eye(eng.n)
    ref = np.linalg.solve(A, -r.cpu().numpy()[:, 0])
    assert int(status[0]) == 0
    assert rel_err(sol.cpu().numpy()[:, 0], ref) < 1e-9 * max(1.0, np.linalg.cond(A) * 1e-6)


def _well_posed_case():
    """Advection-diffusion FVM, dirichlet/neumann data (non-singular blocks)."""
    import splitfxm_b200 as sf
    from splitfxm_b200.models.advection_diffusion import AdvectionDiffusion
    comps = ["u", "v", "w"]
    bcs = {"u": {"left": {"dirichlet": 1.0}, "right": {"dirichlet": 0.0}},
           "v": {"left": {"dirichlet": 3}, "right": {"dirichlet": 4}},
           "w": {"left": {"dirichlet": 2}, "right": {"neumann": 0.0}}}
    ics = {"u": "gaussian", "v": "rarefaction", "w": "tophat"}
    d = sf.Domain.from_size(40, 1, 1, comps)
    s = sf.Simulation(d, AdvectionDiffusion(0.2, 0.05, "FVM"), ics, bcs, sf.FVSchemes.LAX_FRIEDRICHS)
    return s, d, comps, bcs


@pytest.mark.parametrize("split,locs,kw", [
    (False, None, {}),
    (True, [1, 2], {}),
    (True, [1], {"dt0": 0.05, "dtmax": 10.0}),
    (False, None, {"armijo": True}),
    (True, [2], {"bounds": [[-1.0, -1.0, -1.0], [5.0, 5.0, 5.0]]}),
])
def test_steady_state_matches_oracle_spec(split, locs, kw):
    from oracle import core, newton as onewton
    s, d, comps, bcs = _well_posed_case()
    x_all = np.array(d.positions())
    P = core.Problem(core.AdvectionDiffusion(0.2, 0.05), "FVM", "LAX_FRIEDRICHS", x_all[1:-1], 1, 1, bcs, comps)
    x0 = d.listify_interior(split, locs)
    f = lambda u: P.residual(u, split, locs)
    jac = lambda u: P.jacobian_matrix(u, split, locs)
    okw = dict(kw)
    if "bounds" in okw:
        okw["bounds"] = s.extend_bounds(okw["bounds"], 40, 3, split, locs)
    if split:
        xf, _, it_ref = onewton.split_newton(f, jac, x0, [40 * i for i in locs], sparse=True, **okw)
    else:
        xf, _, it_ref = onewton.newton(f, jac, x0, sparse=True, **okw)
    it = s.steady_state(split=split, split_locs=locs, **kw)
    assert it == it_ref
    got = d.listify_interior(split, locs)
    assert rel_err(got, xf) < 1e-10
    # and it is a steady state (the pseudo-transient runs stop on the step-size criterion earlier)
    if "dt0" not in kw:
        assert np.max(np.abs(s.get_residuals_from_list(got, split, locs))) < 1e-6


def test_ensemble_newton_members_match_single():
    import splitfxm_b200 as sf
    from splitfxm_b200.ensemble import EnsembleSimulation
    from splitfxm_b200.models.advection_diffusion import AdvectionDiffusion
    s, d, comps, bcs = _well_posed_case()
    B = 40
    cs = np.linspace(0.05, 0.5, B)
    nus = np.logspace(-2, -1, B)
    ens = EnsembleSimulation(d, "ad", {"c": cs, "nu": nus}, bcs, sf.FVSchemes.LAX_FRIEDRICHS, method="FVM")
    x0 = d.listify_interior(True, [1, 2])
    Y, its = ens.steady_state(np.repeat(x0[:, None], B, axis=1), split=True, split_locs=[1, 2])
    Y = Y.cpu().numpy()
    for k in (0, 17, 39):
        d1 = sf.Domain.from_arrays(d._x, d._U, 1, 1, comps)
        s1 = sf.Simulation(d1, AdvectionDiffusion(cs[k], nus[k], "FVM"), {}, bcs, sf.FVSchemes.LAX_FRIEDRICHS)
        it1 = s1.steady_state(split=True, split_locs=[1, 2])
        assert it1 == int(its[k])
        assert np.array_equal(d1.listify_interior(True, [1, 2]), Y[:, k])


@pytest.mark.parametrize("bcs_kind", ["dirichlet", "periodic"])
def test_pcr_large_single_mesh(bcs_kind):
    """nx = 200,000 cells, one mesh (batch 1): the PCR solve satisfies (J + shift I) s = -f to round-off, checked with
    a banded mat-vec on the host (size-independent property; the dense reference is out of reach at this size)."""
    import torch
    import splitfxm_b200 as sf
    from splitfxm_b200.models.advection_diffusion import AdvectionDiffusion
    nx = 200_000
    comps = ["u", "v", "w"]
    if bcs_kind == "dirichlet":
        bcs = {c: {"left": {"dirichlet": 1.0}, "right": {"dirichlet": 0.0}} for c in comps}
    else:
        bcs = {"u": {"left": "periodic", "right": "periodic"}, "v": {"left": {"dirichlet": 3}, "right": "periodic"},
               "w": {"left": {"neumann": 0.0}, "right": {"dirichlet": 2}}}
    d = sf.Domain.from_size(nx, 1, 1, comps)
    s = sf.Simulation(d, AdvectionDiffusion(0.2, 0.01, "FVM"), {"u": "gaussian", "v": "sine", "w": "tophat"}, bcs,
                      sf.FVSchemes.LAX_FRIEDRICHS)
    eng = s._engine()
    y = eng.to_device(d.listify_interior(False))
    r = torch.empty_like(y)
    J = eng.jacobian(y, r0=r)
    shift = torch.full((1,), -1.0e4, dtype=torch.float64, device=y.device)     # pseudo-transient term, 1/dt = 1e4
    sol = torch.zeros_like(y)
    status = torch.zeros(1, dtype=torch.int32, device=y.device)
    eng.solve(J, r, sol, shift=shift, status=status, method="pcr")
    assert int(status[0]) == 0
    Jh = J.cpu().numpy()[..., 0]                      # [nx, 3, nv, nv]
    x = sol.cpu().numpy()[:, 0].reshape(nx, 3)
    Ax = -1.0e4 * x
    for dslot in range(3):
        xn = np.roll(x, -(dslot - 1), axis=0)         # x at cell (i + d - 1) mod nx (wrap slots are zero unless periodic)
        Ax += np.einsum("irc,ic->ir", Jh[:, dslot], xn)
    f = r.cpu().numpy()[:, 0].reshape(nx, 3)
    assert np.max(np.abs(Ax + f)) <= 1e-9 * max(1.0, np.max(np.abs(f)))


@pytest.mark.parametrize("name,split,locs", [
    ("ad_fdm_rbias_nx40_mixedbc", False, None),       # nb (1, 2), w left-periodic, 3x3 blocks -> 6x6 super-blocks
    ("ad_fdm_rbias_nx40_mixedbc", True, [1, 2]),
    ("ad_fdm_rbias_nx40_nb21", False, None),          # nb (2, 1)
    ("ad_fdm_rbias_nx250_split1", True, [1]),         # README BCs: u periodic on both sides
])
def test_band4_solve_right_bias(name, split, locs):
    import torch
    from oracle import core
    z = load(name)
    s, d, m = product_simulation(z)
    eng = s._engine()
    y_host = core.to_flat(core.to_block(z["x0"], m["nx"], len(m["comps"]), m["split"], m["locs"]), split, locs)
    y = eng.to_device(y_host)
    r = torch.empty_like(y)
    J = eng.jacobian(y, split, locs, r0=r)
    shift = torch.full((1,), -50.0, dtype=torch.float64, device=y.device)
    sol = torch.zeros_like(y)
    status = torch.zeros(1, dtype=torch.int32, device=y.device)
    eng.solve(J, r, sol, split, locs, shift=shift, status=status)
    A = _dense_from_banded(eng, J.cpu().numpy()[..., 0], split, locs) - 50.0 * np.eye(eng.n)
    ref = np.linalg.solve(A, -r.cpu().numpy()[:, 0])
    assert int(status[0]) == 0
    assert rel_err(sol.cpu().numpy()[:, 0], ref) < 1e-9 * max(1.0, np.linalg.cond(A) * 1e-6)


def test_steady_state_right_bias_matches_oracle_spec():
    """Newton with the 4-cell RIGHT_BIAS stencil (odd mesh, no periodic variable): same iterates and iteration count
    as the specification in oracle/newton.py."""
    import splitfxm_b200 as sf
    from splitfxm_b200.models.advection_diffusion import AdvectionDiffusion
    from oracle import core, newton as onewton
    comps = ["u", "v"]
    bcs = {"u": {"left": {"dirichlet": 1.0}, "right": {"dirichlet": 0.0}},
           "v": {"left": {"dirichlet": 3}, "right": {"neumann": 0.0}}}
    d = sf.Domain.from_size(41, 1, 2, comps)
    s = sf.Simulation(d, AdvectionDiffusion(0.2, 0.05, "FDM"), {"u": "gaussian", "v": "rarefaction"}, bcs, sf.FDSchemes.RIGHT_BIAS)
    P = core.Problem(core.AdvectionDiffusion(0.2, 0.05), "FDM", "RIGHT_BIAS", np.array(d.positions(interior=True)), 1, 2, bcs, comps)
    x0 = d.listify_interior(True, [1])
    xf, _, it_ref = onewton.split_newton(lambda u: P.residual(u, True, [1]), lambda u: P.jacobian_matrix(u, True, [1]), x0,
                                         [41], sparse=True, dt0=0.05, dtmax=10.0)
    it = s.steady_state(split=True, split_locs=[1], dt0=0.05, dtmax=10.0)
    assert it == it_ref
    assert rel_err(d.listify_interior(True, [1]), xf) < 1e-9


@pytest.mark.parametrize("case,split,locs", [
    ("ad_dirichlet", True, [1]),        # plain Thomas, 1x1 and 2x2 segments
    ("ad_dirichlet", False, None),      # 3x3 blocks
    ("ad_bench", True, [1]),            # bench BCs: (v, w) segment with a lower-left corner block only
    ("ad_readme", True, [1, 2]),        # u fully cyclic (condensation), v plain, w right-periodic
    ("ad_readme", False, None),
    ("ad_fvm_lf", True, [2]),
    ("sod_cons", False, None),          # dense Euler blocks
])
def test_fused_newton_solve_bit_identical_to_assemble_then_solve(case, split, locs):
    """sfxm_newton_solve_fused (Jacobian block rows evaluated inside the Thomas sweep) against sfxm_jacobian_segments +
    sfxm_block_tridiag_solve on the same ensemble: same kernels, same operation order -> identical bits."""
    import torch
    import splitfxm_b200 as sf
    from splitfxm_b200.ensemble import EnsembleSimulation
    B, nx = 64, 37
    rng = np.random.default_rng(3)
    readme = {"u": {"left": "periodic", "right": "periodic"}, "v": {"left": {"dirichlet": 3}, "right": {"dirichlet": 4}},
              "w": {"left": {"dirichlet": 2}, "right": "periodic"}}
    if case == "sod_cons":
        comps = ["rho", "rhou", "E"]
        bcs = {n: {"left": {"neumann": 0.0}, "right": {"neumann": 0.0}} for n in comps}
        d = sf.Domain.from_size(nx, 1, 1, comps)
        ens = EnsembleSimulation(d, "euler_cons", {"gamma": np.linspace(1.3, 1.5, B)}, bcs, sf.FVSchemes.LAX_FRIEDRICHS, "FVM")
        U = np.stack([1.0 + 0.1 * rng.random((nx, B)), 0.1 * rng.standard_normal((nx, B)), 2.5 + 0.1 * rng.random((nx, B))], axis=1)
    else:
        comps = ["u", "v", "w"]
        bcs = {"ad_dirichlet": {c: {"left": {"dirichlet": 1.0}, "right": {"neumann": 0.5}} for c in comps},
               "ad_bench": dict(readme, u={"left": {"dirichlet": 0.0}, "right": {"dirichlet": 0.0}}),
               "ad_readme": readme, "ad_fvm_lf": readme}[case]
        d = sf.Domain.from_size(nx, 1, 1, comps)
        scheme, method = (sf.FVSchemes.LAX_FRIEDRICHS, "FVM") if case == "ad_fvm_lf" else (sf.FDSchemes.CENTRAL, "FDM")
        ens = EnsembleSimulation(d, "ad", {"c": np.linspace(0.05, 0.5, B), "nu": np.logspace(-3, -1, B)}, bcs, scheme, method)
        U = rng.standard_normal((nx, 3, B))
    eng = ens.eng
    from splitfxm_b200.engine import flat_index_table, split_segments
    idx = flat_index_table(nx, 3, split_segments(3, split, locs))
    Y = np.empty((3 * nx, B))
    Y[idx.reshape(-1)] = U.reshape(3 * nx, B)
    y = eng.to_device(Y)
    _, segs = eng.problem(split, locs)
    shift = torch.linspace(-30.0, -60.0, B, dtype=torch.float64, device=y.device)
    active = torch.ones(B, dtype=torch.int32, device=y.device)
    active[5] = 0
    for k in range(len(segs)):
        r_a, r_b = torch.zeros_like(y), torch.zeros_like(y)
        s_a, s_b = torch.zeros_like(y), torch.zeros_like(y)
        st_a = torch.zeros(B, dtype=torch.int32, device=y.device)
        st_b = torch.zeros(B, dtype=torch.int32, device=y.device)
        J = eng.jacobian(y, split, locs, 1e-8, r0=r_a, seg_lo=k, seg_hi=k + 1)
        eng.solve(J, r_a, s_a, split, locs, k, k + 1, shift=shift, active=active, status=st_a, method="thomas")
        eng.newton_solve_fused(y, s_b, split, locs, k, 1e-8, shift=shift, active=active, status=st_b, r0=r_b)
        assert int(st_a.sum()) == 0 and int(st_b.sum()) == 0
        assert torch.equal(s_a, s_b), f"segment {k}: max diff {float((s_a - s_b).abs().max())}"
        assert float(s_b[:, 5].abs().max()) == 0.0                      # masked member untouched
        act = active.bool()
        lo, hi = segs[k][0] * nx, (segs[k][0] + segs[k][1]) * nx          # the segment's rows of f(y)
        assert torch.equal(r_a[lo:hi][:, act], r_b[lo:hi][:, act])
        assert float(s_b.abs().max()) > 0.0


def test_ensemble_newton_fused_equals_unfused():
    import torch
    import splitfxm_b200 as sf
    from splitfxm_b200.ensemble import EnsembleSimulation
    from splitfxm_b200.newton import NewtonDriver
    s, d, comps, bcs = _well_posed_case()
    B = 2048
    ens = EnsembleSimulation(d, "ad", {"c": np.linspace(0.05, 0.5, B), "nu": np.logspace(-2, -1, B)}, bcs,
                             sf.FVSchemes.LAX_FRIEDRICHS, method="FVM")
    x0 = d.listify_interior(True, [1])
    Y0 = ens.eng.to_device(np.repeat(x0[:, None], B, axis=1))
    assert ens.eng.fused_solve_supported(True, [1])
    for kw in ({}, {"dt0": 0.05, "dtmax": 10.0}, {"armijo": True}):
        ya, yb = Y0.clone(), Y0.clone()
        drv_a = NewtonDriver(ens.eng, True, [1], fused=True, **kw)
        drv_b = NewtonDriver(ens.eng, True, [1], fused=False, **kw)
        assert drv_a.fused and not drv_b.fused
        ita, itb = drv_a.run(ya), drv_b.run(yb)
        assert torch.equal(ita, itb) and torch.equal(ya, yb)


@pytest.mark.parametrize("name,split,locs", [
    ("ad_fvm_lf_nx40_mixedbc", False, None),
    ("ad_fvm_lf_nonuniform", True, [1, 2]),
    ("ad_fdm_central_nx250_split1", True, [1]),
    ("sod_cons_nx100_t0p05", False, None),
])
def test_pcr_one_launch_shared_memory_equals_multi_launch(name, split, locs, monkeypatch):
    """Small meshes take the one-launch PCR with the rows in shared memory; it runs the same row arithmetic as the
    multi-launch passes, so the two agree bit for bit (batch 1 and a small ensemble)."""
    import torch
    from oracle import core
    z = load(name)
    s, d, m = product_simulation(z)
    eng = s._engine()
    y = eng.to_device(core.to_flat(core.to_block(z["x0"], m["nx"], len(m["comps"]), m["split"], m["locs"]), split, locs))
    r = torch.empty_like(y)
    J = eng.jacobian(y, split, locs, r0=r)
    shift = torch.full((1,), -3.0, dtype=torch.float64, device=y.device)
    out = {}
    for flag in ("0", "1"):
        monkeypatch.setenv("SFXM_PCR_SMEM", flag)
        sol = torch.zeros_like(y)
        status = torch.zeros(1, dtype=torch.int32, device=y.device)
        from splitfxm_b200 import _lib
        before = _lib.LAUNCHES[0]
        eng.solve(J, r, sol, split, locs, shift=shift, status=status, method="pcr")
        out[flag] = (sol, _lib.LAUNCHES[0] - before)
        assert int(status[0]) == 0
    assert torch.equal(out["0"][0], out["1"][0])
    _, segs = eng.problem(split, locs)
    assert out["1"][1] == len(segs) and out["0"][1] > 5 * len(segs)


@pytest.mark.parametrize("nx", [2048, 2049, 2305, 70_000, 600_000])
@pytest.mark.parametrize("bcs_kind,split,locs", [("dirichlet", False, None), ("periodic", False, None),
                                                 ("periodic", True, [1]), ("periodic", True, [1, 2])])
def test_chunked_pcr_matches_plain_pcr(nx, bcs_kind, split, locs, monkeypatch):
    """Large meshes: nested dissection around the shared-memory PCR (separators every 256 cells, incl. meshes whose last
    cell is itself a separator and two-level recursion) against the plain multi-pass PCR, and both against the banded
    mat-vec residual."""
    import torch
    import splitfxm_b200 as sf
    from splitfxm_b200.models.advection_diffusion import AdvectionDiffusion
    comps = ["u", "v", "w"]
    if bcs_kind == "dirichlet":
        bcs = {c: {"left": {"dirichlet": 1.0}, "right": {"neumann": 0.5}} for c in comps}
    else:
        bcs = {"u": {"left": "periodic", "right": "periodic"}, "v": {"left": {"dirichlet": 3}, "right": "periodic"},
               "w": {"left": {"neumann": 0.0}, "right": {"dirichlet": 2}}}
    d = sf.Domain.from_size(nx, 1, 1, comps)
    s = sf.Simulation(d, AdvectionDiffusion(0.2, 0.01, "FVM"), {"u": "gaussian", "v": "sine", "w": "tophat"}, bcs,
                      sf.FVSchemes.LAX_FRIEDRICHS)
    eng = s._engine()
    y = eng.to_device(d.listify_interior(split, locs))
    r = torch.empty_like(y)
    J = eng.jacobian(y, split, locs, r0=r)
    shift = torch.full((1,), -float(nx) ** 2 * 0.05, dtype=torch.float64, device=y.device)
    sols = {}
    for flag in ("0", "1"):
        monkeypatch.setenv("SFXM_PCR_CHUNKED", flag)
        sol = torch.zeros_like(y)
        status = torch.zeros(1, dtype=torch.int32, device=y.device)
        eng.solve(J, r, sol, split, locs, shift=shift, status=status, method="pcr")
        assert int(status[0]) == 0
        sols[flag] = sol.cpu().numpy()[:, 0]
    scale = np.max(np.abs(sols["0"]))
    assert np.max(np.abs(sols["1"] - sols["0"])) <= 1e-11 * scale
    # residual of the chunked solution through the banded mat-vec (same-segment entries only)
    from oracle import core
    Jh = J.cpu().numpy()[..., 0]
    x = core.to_block(sols["1"], nx, 3, split, locs)
    f = core.to_block(r.cpu().numpy()[:, 0], nx, 3, split, locs)
    _, segs = eng.problem(split, locs)
    same = np.zeros((3, 3), dtype=bool)
    for (o, sz) in segs:
        same[o:o + sz, o:o + sz] = True
    Ax = float(shift[0]) * x
    for dslot in range(3):
        xn = np.roll(x, -(dslot - 1), axis=0)
        Ax += np.einsum("irc,ic->ir", Jh[:, dslot] * same[None], xn)
    assert np.max(np.abs(Ax + f)) <= 1e-9 * max(1.0, np.max(np.abs(f)))


@pytest.mark.parametrize("B,n", [(1, 777), (5, 300), (96, 1201)])
def test_update_norms_equals_axpy_then_sumsq(B, n):
    """sfxm_update_norms = sfxm_axpy_masked + two sfxm_sumsq, bit for bit (same tiles, same order of additions)."""
    import torch
    from splitfxm_b200.engine import Engine
    from splitfxm_b200 import _lib
    eng = Engine(_lib.MODEL_AD, _lib.FD_CENTRAL, 3, 1, 1, np.linspace(0, 1, n // 3 if n % 3 == 0 else 100), 0.0, 1.0,
                 np.array([0.2, 0.01]), batch=B)
    eng.n = n                                           # only the vector kernels are used here
    g = torch.Generator(device="cuda").manual_seed(B)
    y = torch.randn((n, B), dtype=torch.float64, device="cuda", generator=g)
    s = torch.randn((n, B), dtype=torch.float64, device="cuda", generator=g)
    active = (torch.arange(B, device="cuda") % 3 != 1).to(torch.int32)
    lo, hi = 17, n - 5
    y_ref = y.clone()
    eng.axpy(y_ref, s, None, active, lo, hi)
    ss_ref, yy_ref = eng.sumsq(s, lo, hi).clone(), eng.sumsq(y_ref, lo, hi).clone()
    y_new = y.clone()
    ss, yy = eng.update_norms(y_new, s, active, lo, hi)
    assert torch.equal(y_new, y_ref) and torch.equal(ss, ss_ref) and torch.equal(yy, yy_ref)
    y_keep = y.clone()
    ss2, yy2 = eng.update_norms(y_keep, s, None, lo, hi, update=False)
    assert torch.equal(y_keep, y) and torch.equal(ss2, ss_ref) and torch.equal(yy2, eng.sumsq(y, lo, hi))
