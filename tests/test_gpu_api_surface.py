"""The reference's API outside Simulation, against vectors produced by the unmodified reference
(tests/golden/make_api_golden.py -> api_surface.npz): System.residuals(d) (system.py:28-70), the equation objects'
residuals(cell_sub, scheme, scheme_opts) (equations/fv_transport.py:27-59, fd_transport.py:26-51), and
Refiner.perform_changes(d, loc, keep) (refine.py:239-282).  Advection-diffusion: bit-exact; Euler: 1e-10."""
import json
import os

import numpy as np
import pytest

from helpers import GOLDEN, rel_err

pytestmark = pytest.mark.gpu
Z = np.load(os.path.join(GOLDEN, "api_surface.npz"))
CASES = json.loads(str(Z["cases"]))
CHANGES = json.loads(str(Z["changes"]))


def _model(c):
    from splitfxm_b200.models.advection_diffusion import AdvectionDiffusion
    from splitfxm_b200.models.euler1d import Euler1DConservative
    return AdvectionDiffusion(c["c"], c["nu"], c["method"]) if c["model"] == "ad" else Euler1DConservative(c["gamma"], c["method"])


@pytest.mark.parametrize("c", CASES, ids=[c["tag"] for c in CASES])
def test_system_residuals(c):
    import splitfxm_b200 as sf
    from splitfxm_b200.system import System
    d = sf.Domain.from_arrays(Z[c["tag"] + "_x"], Z[c["tag"] + "_U"], c["nbl"], c["nbr"], c["comps"])
    scheme = (sf.FVSchemes if c["method"] == "FVM" else sf.FDSchemes)[c["scheme"]]
    r = System(_model(c), scheme, {}).residuals(d)
    ref = Z[c["tag"] + "_res"]
    assert r.shape == ref.shape == (c["nx"], len(c["comps"]))
    if c["model"] == "ad":
        assert np.array_equal(r, ref)
    else:
        assert rel_err(r, ref) < 1e-10


@pytest.mark.parametrize("c", CASES, ids=[c["tag"] for c in CASES])
def test_equation_residuals_on_one_stencil(c):
    """The Python callables F, D, S, dFdU of the model's equation object evaluated on the host, the stencil arithmetic
    in one kernel launch - also what a model without a device functor gets from System.residuals."""
    import splitfxm_b200 as sf
    d = sf.Domain.from_arrays(Z[c["tag"] + "_x"], Z[c["tag"] + "_U"], c["nbl"], c["nbr"], c["comps"])
    scheme = (sf.FVSchemes if c["method"] == "FVM" else sf.FDSchemes)[c["scheme"]]
    i = d.ilo() + c["cell"]
    cell_sub = [d.cells()[i + o] for o in range(-c["nbl"], c["nbr"] + 1)]
    r = _model(c).equation().residuals(cell_sub, scheme, {"limiter": None})
    ref = Z[c["tag"] + "_cell"]
    if c["model"] == "ad":
        assert np.array_equal(r, ref)
    else:
        assert rel_err(r, ref) < 1e-10
    with pytest.raises(sf.SFXM):
        _model(c).equation().residuals(cell_sub[:-1], scheme)


def test_system_residuals_of_an_opaque_model_loops_over_cells():
    """A model with Python callables only (no device functor): System.residuals walks the cells like the reference and
    each cell's arithmetic is a kernel launch; same numbers as the device functor of the same equations."""
    import splitfxm_b200 as sf
    from splitfxm_b200.equations.fv_transport import FVTransportEquation
    from splitfxm_b200.system import System
    c = CASES[0]

    class Opaque(sf.Model):
        def __init__(self):
            self._equation = FVTransportEquation(lambda u: c["c"] * np.asarray(u), lambda u: c["nu"] * np.asarray(u),
                                                 lambda u: np.zeros(1), lambda u: c["c"] * np.eye(len(u)))
            self._device = None

    d = sf.Domain.from_arrays(Z[c["tag"] + "_x"], Z[c["tag"] + "_U"], c["nbl"], c["nbr"], c["comps"])
    r = System(Opaque(), sf.FVSchemes.LAX_FRIEDRICHS, {}).residuals(d)
    assert np.array_equal(r, Z[c["tag"] + "_res"])


@pytest.mark.parametrize("c", CHANGES, ids=[c["tag"] for c in CHANGES])
def test_perform_changes(c):
    import splitfxm_b200 as sf
    d = sf.Domain.from_arrays(Z[c["tag"] + "_x0"], Z[c["tag"] + "_U0"], 1, 1, ["u", "v"])
    loc = {int(k): v for k, v in c["loc"].items()}
    keep = {int(k): v for k, v in c["keep"].items()}
    sf.Refiner().perform_changes(d, loc, keep)
    assert np.array_equal(np.array(d.positions()), Z[c["tag"] + "_x1"])
    assert np.array_equal(d.values(), Z[c["tag"] + "_U1"])
    assert len(d.interior()) == Z[c["tag"] + "_x1"].shape[0] - 2


def test_domain_stays_on_the_device_between_solver_calls():
    """evolve -> refine -> evolve -> steady_state without a host access in between: the Domain's host copy is stale
    throughout (nothing was downloaded), and the first host access afterwards sees the final state."""
    import contextlib
    import io
    import splitfxm_b200 as sf
    from splitfxm_b200.models.advection_diffusion import AdvectionDiffusion
    comps = ["u", "v", "w"]
    bcs = {"u": {"left": {"dirichlet": 1.0}, "right": {"dirichlet": 0.0}}, "v": {"left": {"dirichlet": 3}, "right": {"dirichlet": 4}},
           "w": {"left": {"dirichlet": 2}, "right": {"neumann": 0.0}}}
    d = sf.Domain.from_size(200, 1, 1, comps)
    s = sf.Simulation(d, AdvectionDiffusion(0.2, 0.05, "FVM"), {"u": "gaussian", "v": "rarefaction", "w": "tophat"}, bcs,
                      sf.FVSchemes.LAX_FRIEDRICHS)
    s._r.set_criteria(3.0, 0.05, 0.05, -0.001)
    s.evolve(0.01)
    assert not d._host_ok and d._dev_ok
    with contextlib.redirect_stdout(io.StringIO()):
        s._r.refine(d)
    assert not d._host_ok and d._nx > 200
    s.evolve(0.002, method="Euler", max_step=2e-5)      # explicit: nu dt / dx^2 = 0.16 on the refined mesh
    it = s.steady_state(split=True, split_locs=[1, 2])
    assert not d._host_ok and it >= 1
    # same sequence with a host access after every call (forces download + re-upload): identical result
    d2 = sf.Domain.from_size(200, 1, 1, comps)
    s2 = sf.Simulation(d2, AdvectionDiffusion(0.2, 0.05, "FVM"), {"u": "gaussian", "v": "rarefaction", "w": "tophat"}, bcs,
                       sf.FVSchemes.LAX_FRIEDRICHS)
    s2._r.set_criteria(3.0, 0.05, 0.05, -0.001)
    s2.evolve(0.01)
    d2.interior_U()
    with contextlib.redirect_stdout(io.StringIO()):
        s2._r.refine(d2)
    d2.interior_x()
    s2.evolve(0.002, method="Euler", max_step=2e-5)
    d2.values()
    assert s2.steady_state(split=True, split_locs=[1, 2]) == it
    assert np.isfinite(d.values()).all()
    assert np.array_equal(np.array(d.positions()), np.array(d2.positions())) and np.array_equal(d.values(), d2.values())


def test_model_parameters_are_read_at_call_time():
    """the reference's F / D read self.c / self.nu when called; changing them after construction must count."""
    import splitfxm_b200 as sf
    from splitfxm_b200.models.advection_diffusion import AdvectionDiffusion
    m = AdvectionDiffusion(0.2, 0.01, "FDM")
    d = sf.Domain.from_size(30, 1, 1, ["u"])
    s = sf.Simulation(d, m, {"u": "sine"}, {"u": {"left": "periodic", "right": "periodic"}}, sf.FDSchemes.CENTRAL)
    x0 = d.listify_interior(False)
    r1 = s.get_residuals_from_list(x0)
    m.c, m.nu = 0.4, 0.02
    r2 = s.get_residuals_from_list(x0)
    d3 = sf.Domain.from_size(30, 1, 1, ["u"])
    s3 = sf.Simulation(d3, AdvectionDiffusion(0.4, 0.02, "FDM"), {"u": "sine"}, {"u": {"left": "periodic", "right": "periodic"}},
                       sf.FDSchemes.CENTRAL)
    assert not np.array_equal(r1, r2) and np.array_equal(r2, s3.get_residuals_from_list(x0))
