"""Slab decomposition (BASELINE config 5) emulated on one GPU: G slabs with local halo copies must reproduce the
single-mesh residual and explicit Euler evolution bit for bit."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

README_BCS = {"u": {"left": "periodic", "right": "periodic"},
              "v": {"left": {"dirichlet": 3}, "right": {"dirichlet": 4}},
              "w": {"left": {"dirichlet": 2}, "right": "periodic"}}
NEUMANN0 = {n: {"left": {"neumann": 0.0}, "right": {"neumann": 0.0}} for n in ("rho", "rhou", "E")}


def _slabs(x_global, comps, model, params, bcs, scheme, method, G, nbl=1, nbr=1):
    from splitfxm_b200.distributed import LocalTransport, SlabSimulation
    tr = LocalTransport()
    for r in range(G):
        tr.slabs.append(SlabSimulation(x_global, comps, model, params, bcs, scheme, method, r, G, nbl=nbl, nbr=nbr,
                                       transport=tr))
    return tr


def _split_state(tr, U_global, split, locs):
    """Global [nx, nv] block -> per-slab flat device vectors."""
    from oracle import core
    ys = []
    for s in tr.slabs:
        ys.append(s.eng.to_device(core.to_flat(U_global[s.lo:s.hi], split, locs)))
    return ys


def _gather(tr, vecs, split, locs):
    from oracle import core
    return np.concatenate([core.to_block(v.cpu().numpy()[:, 0], s.hi - s.lo, s.nv, split, locs)
                           for s, v in zip(tr.slabs, vecs)])


@pytest.mark.parametrize("G", [2, 3])
@pytest.mark.parametrize("case", ["ad_fvm_readme", "ad_fdm_rbias", "sod_cons", "ad_one_sided_periodic"])
def test_slabs_reproduce_single_mesh(case, G):
    import splitfxm_b200 as sf
    from splitfxm_b200.models.advection_diffusion import AdvectionDiffusion
    from splitfxm_b200.models.euler1d import Euler1DConservative
    rng = np.random.default_rng(7)
    nx = 64
    if case == "ad_fvm_readme":
        comps, bcs, nbl, nbr = ["u", "v", "w"], README_BCS, 1, 1
        model, mname, params, scheme, method = AdvectionDiffusion(0.2, 0.001, "FVM"), "ad", {"c": 0.2, "nu": 0.001}, sf.FVSchemes.LAX_FRIEDRICHS, "FVM"
        split, locs = True, [1, 2]
        U = rng.standard_normal((nx, 3))
    elif case == "ad_one_sided_periodic":          # bc.py allows periodic on one side only
        comps, nbl, nbr = ["u", "v", "w"], 1, 1
        bcs = {"u": {"left": "periodic", "right": "outflow"}, "v": {"left": {"dirichlet": 3}, "right": {"neumann": 0.5}},
               "w": {"left": "outflow", "right": {"dirichlet": 1.0}}}
        model, mname, params, scheme, method = AdvectionDiffusion(0.2, 0.01, "FVM"), "ad", {"c": 0.2, "nu": 0.01}, sf.FVSchemes.LAX_FRIEDRICHS, "FVM"
        split, locs = True, [1]
        U = rng.standard_normal((nx, 3))
    elif case == "ad_fdm_rbias":
        comps, bcs, nbl, nbr = ["u", "v", "w"], README_BCS, 1, 2
        model, mname, params, scheme, method = AdvectionDiffusion(-0.3, 0.01, "FDM"), "ad", {"c": -0.3, "nu": 0.01}, sf.FDSchemes.RIGHT_BIAS, "FDM"
        split, locs = False, None
        U = rng.standard_normal((nx, 3))
    else:
        comps, bcs, nbl, nbr = ["rho", "rhou", "E"], NEUMANN0, 1, 1
        model, mname, params, scheme, method = Euler1DConservative(1.4, "FVM"), "euler_cons", {"gamma": 1.4}, sf.FVSchemes.LAX_FRIEDRICHS, "FVM"
        split, locs = True, [1]
        U = np.stack([1.0 + 0.1 * rng.random(nx), 0.1 * rng.standard_normal(nx), 2.5 + 0.1 * rng.random(nx)], axis=1)
    d = sf.Domain.from_size(nx, nbl, nbr, comps)
    d.interior_U()[:, :] = U
    sim = sf.Simulation(d, model, {}, bcs, scheme)
    from oracle import core
    r_ref = sim.get_residuals_from_list(core.to_flat(U, split, locs), split, locs)
    r_ref = core.to_block(r_ref, nx, 3, split, locs)
    x_global = np.array(d.positions(interior=True))
    tr = _slabs(x_global, comps, mname, params, bcs, scheme, method, G, nbl, nbr)
    ys = _split_state(tr, U, split, locs)
    tr.exchange_all([s.pack(y, split, locs) for s, y in zip(tr.slabs, ys)])
    rs = [s.residual(y, None, split, locs, exchange=False) for s, y in zip(tr.slabs, ys)]
    assert np.array_equal(_gather(tr, rs, split, locs), r_ref)


def test_slab_euler_evolution_matches_single_mesh():
    import torch
    import splitfxm_b200 as sf
    from splitfxm_b200.models.advection_diffusion import AdvectionDiffusion
    nx, G = 90, 3
    comps = ["u", "v", "w"]
    d = sf.Domain.from_size(nx, 1, 1, comps)
    sim = sf.Simulation(d, AdvectionDiffusion(0.2, 0.001, "FVM"), {"u": "gaussian", "v": "rarefaction", "w": "tophat"},
                        README_BCS, sf.FVSchemes.LAX_FRIEDRICHS)
    U0 = d.values(interior=True)
    x_global = np.array(d.positions(interior=True))
    sim.evolve(0.3, method="Euler", max_step=0.004)
    U_ref = d.values(interior=True)
    tr = _slabs(x_global, comps, "ad", {"c": 0.2, "nu": 0.001}, README_BCS, sf.FVSchemes.LAX_FRIEDRICHS, "FVM", G)
    ys = _split_state(tr, U0, False, None)
    others = [torch.empty_like(y) for y in ys]
    t = 0.0
    while t < 0.3:                                    # same loop as SlabSimulation.evolve_euler, slabs in lock-step
        dt = min(0.3 - t, 0.004)
        tr.exchange_all([s.pack(y) for s, y in zip(tr.slabs, ys)])
        for k, s in enumerate(tr.slabs):
            s.eng.euler_step(ys[k], others[k], dt)
        ys, others = others, ys
        t += dt
    assert np.array_equal(_gather(tr, ys, False, None), U_ref)
