"""sfxm_newton_run (csrc/sfxm_newton_driver.cu): the Newton / Split-Newton loop nest driven from inside the library must be
the Python driver (splitfxm_b200/newton.py) call for call - same iterates and the same per-member iteration counts, bit for
bit - on one mesh (PCR solves, 3- and 4-cell stencils, cyclic segments), on nonlinear problems that need many iterations,
and on ensembles (fused sweeps; members that need different numbers of iterations).  Against the specification itself
(oracle/newton.py) the native driver is what test_gpu_newton.py / test_gpu_newton_nonlinear.py exercise by default."""
import numpy as np
import pytest

from nonlinear_cases import BURGERS_BCS, BURGERS_COMPS, BURGERS_NX, burgers_initial, burgers_model, product_case

pytestmark = pytest.mark.gpu


def _both(eng, y0, split, locs, maxiter=float("inf")):
    import torch
    from splitfxm_b200.newton import NewtonDriver
    out = []
    for native in (True, False):
        drv = NewtonDriver(eng, split, locs, native=native)
        assert drv.native == native
        y = y0.clone()
        it = drv.run(y, maxiter)
        torch.cuda.synchronize()
        out.append((y.clone(), it.clone(), drv.status.clone()))
    (ya, ia, sa), (yb, ib, sb) = out
    assert torch.equal(ia, ib), (ia, ib)
    assert torch.equal(ya, yb)
    assert torch.equal(sa, sb)
    return ya, ia


def _ad_simulation(method, scheme, nb, bcs=None, nx=40):
    import splitfxm_b200 as sf
    from splitfxm_b200.models.advection_diffusion import AdvectionDiffusion
    comps = ["u", "v", "w"]
    bcs = bcs or {"u": {"left": {"dirichlet": 1.0}, "right": {"dirichlet": 0.0}},
                  "v": {"left": {"dirichlet": 3}, "right": {"dirichlet": 4}},
                  "w": {"left": {"dirichlet": 2}, "right": {"neumann": 0.0}}}
    ics = {"u": "gaussian", "v": "rarefaction", "w": "tophat"}
    d = sf.Domain.from_size(nx, nb[0], nb[1], comps)
    return sf.Simulation(d, AdvectionDiffusion(0.2, 0.05, method), ics, bcs, scheme), d


@pytest.mark.parametrize("split,locs", [(False, None), (True, [1]), (True, [1, 2]), (True, [2])])
def test_one_mesh_three_cell_stencil(split, locs):
    import splitfxm_b200 as sf
    s, d = _ad_simulation("FVM", sf.FVSchemes.LAX_FRIEDRICHS, (1, 1))
    eng = s._engine()
    y, it = _both(eng, s._state(eng, split, locs), split, locs)
    assert int(it[0]) >= 2


def test_one_mesh_cyclic_segments_and_iteration_cap():
    import splitfxm_b200 as sf
    bcs = {"u": {"left": "periodic", "right": "periodic"},
           "v": {"left": {"dirichlet": 3}, "right": {"dirichlet": 4}},
           "w": {"left": {"dirichlet": 2}, "right": "periodic"}}
    s, d = _ad_simulation("FDM", sf.FDSchemes.CENTRAL, (1, 1), bcs, nx=64)
    eng = s._engine()
    _both(eng, s._state(eng, True, [1]), True, [1])
    _, it = _both(eng, s._state(eng, True, [1, 2]), True, [1, 2], maxiter=1)
    assert int(it[0]) == 1


def test_one_mesh_four_cell_stencil():
    import splitfxm_b200 as sf
    s, d = _ad_simulation("FDM", sf.FDSchemes.RIGHT_BIAS, (1, 2))
    eng = s._engine()
    _both(eng, s._state(eng, True, [1]), True, [1])


@pytest.mark.parametrize("name", ["burgers_fvm_lf_plain", "burgers_fvm_lf_split1", "euler_duct_plain"])
def test_nonlinear_cases(name):
    s, d, c = product_case(name)
    eng = s._engine()
    y, it = _both(eng, s._state(eng, c["split"], c["locs"]), c["split"], c["locs"])
    assert int(it[0]) >= c["expect"]["min_it"]


def test_ensemble_fused_sweeps_members_with_different_counts():
    import torch
    import splitfxm_b200 as sf
    from splitfxm_b200.ensemble import EnsembleSimulation
    B = 2048
    d = sf.Domain.from_size(BURGERS_NX, 1, 1, ["u", "v", "w"])
    bcs = {"u": {"left": {"dirichlet": 1.0}, "right": {"dirichlet": 0.0}},
           "v": {"left": {"dirichlet": 3}, "right": {"dirichlet": 4}},
           "w": {"left": {"dirichlet": 2}, "right": "periodic"}}
    cs = np.linspace(0.05, 0.5, B)
    nus = np.logspace(-3, -1, B)
    ens = EnsembleSimulation(d, "ad", {"c": cs, "nu": nus}, bcs, sf.FDSchemes.CENTRAL, method="FDM")
    x = np.array(d.positions(interior=True))
    U0 = np.stack([np.exp(-40 * (x - 0.5) ** 2), 3 + x, 2 + 0.0 * x], axis=1)
    from oracle import core
    y0 = core.to_flat(U0, True, [1])
    Y0 = torch.from_numpy(np.repeat(y0[:, None], B, axis=1).copy()).to(ens.eng.device)
    from splitfxm_b200.newton import NewtonDriver
    assert NewtonDriver(ens.eng, True, [1]).fused
    _both(ens.eng, Y0, True, [1])


def test_pipelined_host_steady_states_equal_the_synchronous_call():
    """EnsembleSimulation.steady_state_host with slots (upload / solve / download on three streams, the next input
    prefetched during the Newton iterations) against the synchronous call, alternating two different inputs."""
    import torch
    import splitfxm_b200 as sf
    from splitfxm_b200.ensemble import EnsembleSimulation
    B = 2048
    d = sf.Domain.from_size(40, 1, 1, ["u", "v", "w"])
    bcs = {"u": {"left": {"dirichlet": 1.0}, "right": {"dirichlet": 0.0}},
           "v": {"left": {"dirichlet": 3}, "right": {"dirichlet": 4}},
           "w": {"left": {"dirichlet": 2}, "right": {"neumann": 0.0}}}
    ens = EnsembleSimulation(d, "ad", {"c": np.linspace(0.05, 0.5, B), "nu": np.logspace(-2, -1, B)}, bcs,
                             sf.FDSchemes.CENTRAL, method="FDM")
    rng = np.random.default_rng(11)
    n = ens.eng.n
    inputs = [torch.from_numpy(rng.random((n, B))).pin_memory() for _ in range(2)]
    refs = []
    for Y in inputs:
        out = torch.empty((n, B), dtype=torch.float64).pin_memory()
        its = torch.empty(B, dtype=torch.int64).pin_memory()
        ens.steady_state_host(Y, out, its, True, [1])
        refs.append((out.clone(), its.clone()))
    outs = [torch.empty((n, B), dtype=torch.float64).pin_memory() for _ in range(2)]
    itss = [torch.empty(B, dtype=torch.int64).pin_memory() for _ in range(2)]
    steps = 5
    for k in range(steps):
        nxt = inputs[(k + 1) % 2] if k + 1 < steps else None
        ens.steady_state_host(inputs[k % 2], outs[k % 2], itss[k % 2], True, [1], slot=k % 2, next_Y_host=nxt)
        if k >= 1:                       # the other slot's result: complete once its download event has passed
            ens._ss_slots[(k - 1) % 2]["free"].synchronize()
            assert torch.equal(outs[(k - 1) % 2], refs[(k - 1) % 2][0])
            assert torch.equal(itss[(k - 1) % 2], refs[(k - 1) % 2][1])
    ens.sync_host()
    assert torch.equal(outs[(steps - 1) % 2], refs[(steps - 1) % 2][0])
    assert torch.equal(itss[(steps - 1) % 2], refs[(steps - 1) % 2][1])
