"""Tiny meshes: the reference has no minimum mesh size - its own tests drive Domain.from_size(5, 1, 1, ...)
through jacobian, evolve and steady_state (tests/test_simulation.py:33-125).  Every mesh from the stencil width up must
give the oracle's residual, Jacobian (values + pattern) and Newton step."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

BCS3 = {"u": {"left": "periodic", "right": "periodic"},
        "v": {"left": {"dirichlet": 3}, "right": {"neumann": 0.5}},
        "w": {"left": "outflow", "right": "periodic"}}


@pytest.mark.parametrize("method,scheme,nb", [("FVM", "LAX_FRIEDRICHS", (1, 1)), ("FVM", "UPWIND", (1, 1)),
                                              ("FDM", "CENTRAL", (1, 1)), ("FDM", "RIGHT_BIAS", (1, 2))])
@pytest.mark.parametrize("split,locs", [(False, None), (True, [1]), (True, [1, 2])])
def test_meshes_from_the_stencil_width_up(method, scheme, nb, split, locs):
    import splitfxm_b200 as sf
    from oracle import core
    from splitfxm_b200.models.advection_diffusion import AdvectionDiffusion
    comps = ["u", "v", "w"]
    W = nb[0] + nb[1] + 1
    rng = np.random.default_rng(4)
    for nx in range(W, 2 * W + 3):
        d = sf.Domain.from_size(nx, nb[0], nb[1], comps)
        sch = (sf.FVSchemes if method == "FVM" else sf.FDSchemes)[scheme]
        s = sf.Simulation(d, AdvectionDiffusion(0.2, 0.01, method), {}, BCS3, sch)
        U = rng.standard_normal((nx, 3))
        x0 = core.to_flat(U, split, locs)
        P = core.Problem(core.AdvectionDiffusion(0.2, 0.01), method, scheme, np.array(d.positions(interior=True)), nb[0],
                         nb[1], BCS3, comps)
        assert np.array_equal(s.get_residuals_from_list(x0, split, locs), P.residual(x0, split, locs)), nx
        J = s.jacobian(x0, split, locs).tocoo()
        order = np.lexsort((J.col, J.row))
        rows, cols, vals = P.jacobian_coo(x0, split, locs)
        assert np.array_equal(J.row[order], rows) and np.array_equal(J.col[order], cols), nx
        assert np.array_equal(J.data[order], vals), nx


def test_reference_simulation_fixture_nx5():
    """tests/test_simulation.py:33-62 of the reference: 5 cells, two variables, FDM CENTRAL - jacobian, evolve (incl. the
    implicit Radau path, split) and steady_state must run; the pseudo-transient Newton step is checked against a dense
    solve of the oracle's matrix."""
    import splitfxm_b200 as sf
    from oracle import core
    from splitfxm_b200.models.advection_diffusion import AdvectionDiffusion
    comps = ["u", "v"]
    bcs = {"u": {"left": "periodic", "right": "periodic"}, "v": {"left": {"dirichlet": 1.0}, "right": {"dirichlet": 2.0}}}
    d = sf.Domain.from_size(5, 1, 1, comps)
    s = sf.Simulation(d, AdvectionDiffusion(0.2, 0.05, "FDM"), {"u": "gaussian", "v": "rarefaction"}, bcs, sf.FDSchemes.CENTRAL)
    x0 = d.listify_interior(True, [1])
    P = core.Problem(core.AdvectionDiffusion(0.2, 0.05), "FDM", "CENTRAL", np.array(d.positions(interior=True)), 1, 1, bcs, comps)
    assert np.array_equal(s.jacobian(x0, True, [1]).toarray(), P.jacobian_matrix(x0, True, [1]).toarray())
    s.evolve(0.01, True, [1], method="Radau")
    s.evolve(0.01)
    assert np.isfinite(d.values()).all()
    d.interior_U()[:, :] = core.to_block(x0, 5, 2, True, [1])
    it = s.steady_state(split=True, split_locs=[1], dt0=0.1, dtmax=1.0)     # u is the singular periodic block: shifted
    assert it >= 0
    # one pseudo-transient Newton step of the well-posed variable against a dense solve
    d2 = sf.Domain.from_size(5, 1, 1, comps)
    s2 = sf.Simulation(d2, AdvectionDiffusion(0.2, 0.05, "FDM"), {"u": "gaussian", "v": "rarefaction"}, bcs, sf.FDSchemes.CENTRAL,
                       ss={"maxiter": 1})
    y0 = d2.listify_interior(False)
    s2.steady_state(dt0=0.1)
    A = P.jacobian_matrix(y0).toarray() - 10.0 * np.eye(10)
    step = np.linalg.solve(A, -P.residual(y0))
    assert np.max(np.abs(d2.listify_interior(False) - (y0 + step))) < 1e-10 * np.max(np.abs(y0 + step))
