"""Newton on a slab-decomposed mesh (splitfxm_b200/spike.py, csrc/sfxm_spike.cu): G slabs run as G threads on one GPU
(ThreadGroup); the partitioned block solve must agree with a dense solve of the single-mesh Jacobian, and the slab
Newton / Split-Newton must take the same iterations as the single-mesh run and reach the same state."""
import numpy as np
import pytest

from helpers import rel_err
from test_gpu_slab import NEUMANN0, README_BCS

pytestmark = pytest.mark.gpu

WELL_POSED = {"u": {"left": {"dirichlet": 1.0}, "right": {"dirichlet": 0.0}},
              "v": {"left": {"dirichlet": 3}, "right": {"dirichlet": 4}},
              "w": {"left": {"dirichlet": 2}, "right": {"neumann": 0.0}}}


def _thread_slab(x_global, comps, mname, params, bcs, scheme, method, rank, G, comm):
    from splitfxm_b200.distributed import SlabSimulation
    from splitfxm_b200.spike import ThreadTransport
    slab = SlabSimulation(x_global, comps, mname, params, bcs, scheme, method, rank, G)
    slab.transport = ThreadTransport(comm, slab.left, slab.right)
    return slab


@pytest.mark.parametrize("G", [2, 3])
@pytest.mark.parametrize("case,split,locs", [
    ("ad_readme", True, [1, 2]),      # u cyclic over the slabs, v plain, w right-periodic
    ("ad_readme", False, None),       # 3x3 blocks, mixed periodic / dirichlet columns in the coupling blocks
    ("ad_readme", True, [1]),
    ("sod_cons", False, None),        # dense Euler blocks, neumann ends
])
def test_slab_solve_matches_dense(case, split, locs, G):
    import torch
    import splitfxm_b200 as sf
    from oracle import core
    from splitfxm_b200.models.advection_diffusion import AdvectionDiffusion
    from splitfxm_b200.models.euler1d import Euler1DConservative
    from splitfxm_b200.spike import SlabNewtonEngine, ThreadGroup
    rng = np.random.default_rng(11)
    nx = 50
    if case == "ad_readme":
        comps, bcs = ["u", "v", "w"], README_BCS
        model, mname, params = AdvectionDiffusion(0.2, 0.01, "FVM"), "ad", {"c": 0.2, "nu": 0.01}
        U = rng.standard_normal((nx, 3))
    else:
        comps, bcs = ["rho", "rhou", "E"], NEUMANN0
        model, mname, params = Euler1DConservative(1.4, "FVM"), "euler_cons", {"gamma": 1.4}
        U = np.stack([1.0 + 0.1 * rng.random(nx), 0.1 * rng.standard_normal(nx), 2.5 + 0.1 * rng.random(nx)], axis=1)
    scheme = sf.FVSchemes.LAX_FRIEDRICHS
    d = sf.Domain.from_size(nx, 1, 1, comps)
    d.interior_U()[:, :] = U
    sim = sf.Simulation(d, model, {}, bcs, scheme)
    x0 = core.to_flat(U, split, locs)
    A = sim.jacobian(x0, split, locs).toarray() - 40.0 * np.eye(3 * nx)
    f = sim.get_residuals_from_list(x0, split, locs)
    ref = core.to_block(np.linalg.solve(A, -f), nx, 3, split, locs)
    x_global = np.array(d.positions(interior=True))

    def body(rank, comm):
        slab = _thread_slab(x_global, comps, mname, params, bcs, scheme, "FVM", rank, G, comm)
        eng = SlabNewtonEngine(slab, comm)
        y = slab.eng.to_device(core.to_flat(U[slab.lo:slab.hi], split, locs))
        r = torch.empty_like(y)
        J = eng.jacobian(y, split, locs, r0=r)
        shift = torch.full((1,), -40.0, dtype=torch.float64, device=y.device)
        status = torch.zeros(1, dtype=torch.int32, device=y.device)
        s = torch.zeros_like(y)
        eng.solve(J, r, s, split, locs, shift=shift, status=status)
        assert int(status[0]) == 0
        return core.to_block(s.cpu().numpy()[:, 0], slab.hi - slab.lo, 3, split, locs)

    got = np.concatenate(ThreadGroup(G).run(body))
    assert rel_err(got, ref) < 1e-9 * max(1.0, np.linalg.cond(A) * 1e-6)


@pytest.mark.parametrize("G", [2, 3])
@pytest.mark.parametrize("split,locs,kw", [
    (False, None, {}),
    (True, [1, 2], {}),
    (True, [1], {"dt0": 0.05, "dtmax": 10.0}),
    (False, None, {"armijo": True}),
    (True, [2], {"bounds": [[-1.0, -1.0, -1.0], [5.0, 5.0, 5.0]]}),
])
def test_slab_newton_matches_single_mesh(split, locs, kw, G):
    import splitfxm_b200 as sf
    from oracle import core
    from splitfxm_b200.models.advection_diffusion import AdvectionDiffusion
    from splitfxm_b200.spike import ThreadGroup
    comps, nx = ["u", "v", "w"], 60
    d = sf.Domain.from_size(nx, 1, 1, comps)
    sim = sf.Simulation(d, AdvectionDiffusion(0.2, 0.05, "FVM"), {"u": "gaussian", "v": "rarefaction", "w": "tophat"},
                        WELL_POSED, sf.FVSchemes.LAX_FRIEDRICHS)
    U0 = d.values(interior=True).copy()
    x_global = np.array(d.positions(interior=True))
    it_ref = sim.steady_state(split=split, split_locs=locs, **kw)
    U_ref = d.values(interior=True)

    def body(rank, comm):
        slab = _thread_slab(x_global, comps, "ad", {"c": 0.2, "nu": 0.05}, WELL_POSED, sf.FVSchemes.LAX_FRIEDRICHS, "FVM",
                            rank, G, comm)
        y = slab.eng.to_device(core.to_flat(U0[slab.lo:slab.hi], split, locs))
        its = slab.steady_state(y, split, locs, comm=comm, **kw)
        return int(its[0]), core.to_block(y.cpu().numpy()[:, 0], slab.hi - slab.lo, 3, split, locs)

    out = ThreadGroup(G).run(body)
    assert all(it == it_ref for it, _ in out)
    assert rel_err(np.concatenate([u for _, u in out]), U_ref) < 1e-10


def test_slab_newton_periodic_pseudo_transient():
    """README boundary conditions (u periodic over the slab ring): pseudo-transient Split-Newton, a fixed number of
    iterations, slabs against the single mesh."""
    import splitfxm_b200 as sf
    from oracle import core
    from splitfxm_b200.models.advection_diffusion import AdvectionDiffusion
    from splitfxm_b200.spike import ThreadGroup
    comps, nx, G = ["u", "v", "w"], 64, 4
    d = sf.Domain.from_size(nx, 1, 1, comps)
    sim = sf.Simulation(d, AdvectionDiffusion(0.2, 0.001, "FVM"), {"u": "gaussian", "v": "rarefaction", "w": "tophat"},
                        README_BCS, sf.FVSchemes.LAX_FRIEDRICHS, ss={"maxiter": 3})
    U0 = d.values(interior=True).copy()
    x_global = np.array(d.positions(interior=True))
    it_ref = sim.steady_state(split=True, split_locs=[1, 2], dt0=0.01, dtmax=1.0)
    U_ref = d.values(interior=True)

    def body(rank, comm):
        slab = _thread_slab(x_global, comps, "ad", {"c": 0.2, "nu": 0.001}, README_BCS, sf.FVSchemes.LAX_FRIEDRICHS, "FVM",
                            rank, G, comm)
        y = slab.eng.to_device(core.to_flat(U0[slab.lo:slab.hi], True, [1, 2]))
        its = slab.steady_state(y, True, [1, 2], dt0=0.01, dtmax=1.0, maxiter=3, comm=comm)
        return int(its[0]), core.to_block(y.cpu().numpy()[:, 0], slab.hi - slab.lo, 3, True, [1, 2])

    out = ThreadGroup(G).run(body)
    assert all(it == it_ref for it, _ in out)
    assert rel_err(np.concatenate([u for _, u in out]), U_ref) < 1e-9


@pytest.mark.parametrize("N,B", [(4, 1), (12, 3), (64, 2), (76, 1)])
def test_dense_solve_kernel(N, B):
    """sfxm_dense_solve (the interface system) against NumPy, incl. systems that need row exchanges."""
    import torch
    from splitfxm_b200 import _lib
    rng = np.random.default_rng(N)
    M = rng.standard_normal((B, N, N))
    M[:, 0, 0] = 0.0                                   # forces a pivot search in the first column
    rhs = rng.standard_normal((B, N))
    aug = torch.from_numpy(np.concatenate([M, rhs[:, :, None]], axis=2)).cuda().contiguous()
    z = torch.empty((B, N), dtype=torch.float64, device="cuda")
    status = torch.zeros(B, dtype=torch.int32, device="cuda")
    _lib.check(_lib.load().sfxm_dense_solve(_lib.ptr(aug), N, B, _lib.ptr(z), _lib.ptr(status), _lib.stream_ptr()))
    ref = np.stack([np.linalg.solve(M[b], rhs[b]) for b in range(B)])
    assert int(status.sum()) == 0
    assert rel_err(z.cpu().numpy(), ref) < 1e-10 * max(1.0, max(np.linalg.cond(M[b]) for b in range(B)) * 1e-4)


@pytest.mark.parametrize("G,nx", [(2, 4100), (3, 7000)])
@pytest.mark.parametrize("bcs,split,locs", [(README_BCS, True, [1, 2]), (README_BCS, False, None), (WELL_POSED, True, [1])])
def test_slab_solve_large_slabs_nested_dissection(bcs, split, locs, G, nx):
    """Slabs of >= 2048 cells take the nested-dissection path (separators every 256 cells + both slab ends); the
    gathered step must equal the single-mesh solve."""
    import torch
    import splitfxm_b200 as sf
    from oracle import core
    from splitfxm_b200.models.advection_diffusion import AdvectionDiffusion
    from splitfxm_b200.spike import SlabNewtonEngine, ThreadGroup
    comps = ["u", "v", "w"]
    rng = np.random.default_rng(5)
    U = rng.standard_normal((nx, 3))
    scheme = sf.FVSchemes.LAX_FRIEDRICHS
    d = sf.Domain.from_size(nx, 1, 1, comps)
    d.interior_U()[:, :] = U
    sim = sf.Simulation(d, AdvectionDiffusion(0.2, 0.01, "FVM"), {}, bcs, scheme)
    e1 = sim._engine()
    y1 = e1.to_device(core.to_flat(U, split, locs))
    r1 = torch.empty_like(y1)
    J1 = e1.jacobian(y1, split, locs, r0=r1)
    shift1 = torch.full((1,), -float(nx) ** 2 * 0.05, dtype=torch.float64, device=y1.device)
    s1 = torch.zeros_like(y1)
    e1.solve(J1, r1, s1, split, locs, shift=shift1, method="pcr")
    ref = core.to_block(s1.cpu().numpy()[:, 0], nx, 3, split, locs)
    x_global = np.array(d.positions(interior=True))

    def body(rank, comm):
        slab = _thread_slab(x_global, comps, "ad", {"c": 0.2, "nu": 0.01}, bcs, scheme, "FVM", rank, G, comm)
        assert slab.eng.nx >= 2048
        eng = SlabNewtonEngine(slab, comm)
        y = slab.eng.to_device(core.to_flat(U[slab.lo:slab.hi], split, locs))
        r = torch.empty_like(y)
        J = eng.jacobian(y, split, locs, r0=r)
        shift = torch.full((1,), -float(nx) ** 2 * 0.05, dtype=torch.float64, device=y.device)
        status = torch.zeros(1, dtype=torch.int32, device=y.device)
        s = torch.zeros_like(y)
        eng.solve(J, r, s, split, locs, shift=shift, status=status)
        assert int(status[0]) == 0
        return core.to_block(s.cpu().numpy()[:, 0], slab.hi - slab.lo, 3, split, locs)

    got = np.concatenate(ThreadGroup(G).run(body))
    assert rel_err(got, ref) < 1e-10
