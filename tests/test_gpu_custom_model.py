"""A custom model (viscous Burgers + a coupled 2-variable system) supplied as CUDA device code, compiled into a
plugin by splitfxm_b200.custom and run through the unchanged Simulation API; checked against the oracle's
residual/Jacobian machinery evaluating the same model in NumPy (oracle/core.py with a test-local model)."""
import numpy as np
import pytest

from nonlinear_cases import BURGERS_SRC as BURGERS, OracleBurgers  # noqa: E402


def test_plugin_compiles_on_cpu():
    """nvcc cross-compiles the plugin without a GPU; registration is host-side only."""
    from splitfxm_b200.custom import compile_model
    dm = compile_model(BURGERS, 2, [0.01], "FVM")
    assert dm.model_id >= 16
    assert compile_model(BURGERS, 2, [0.01], "FVM").model_id == dm.model_id      # cached


@pytest.mark.gpu
@pytest.mark.parametrize("method,scheme_name,nb", [("FVM", "LAX_FRIEDRICHS", (1, 1)), ("FVM", "UPWIND", (1, 1)),
                                                   ("FDM", "CENTRAL", (1, 1)), ("FDM", "RIGHT_BIAS", (1, 2))])
def test_custom_model_matches_numpy_evaluation(method, scheme_name, nb):
    import splitfxm_b200 as sf
    from splitfxm_b200.custom import compile_model
    from splitfxm_b200.equations.fd_transport import FDTransportEquation
    from splitfxm_b200.equations.fv_transport import FVTransportEquation
    from oracle import core
    nu = 0.01
    comps = ["a", "b"]
    bcs = {"a": {"left": "periodic", "right": "periodic"}, "b": {"left": {"dirichlet": 0.5}, "right": "outflow"}}

    class Burgers(sf.Model):
        def __init__(self):
            F = lambda u: 0.5 * np.asarray(u) ** 2
            D = lambda u: nu * np.asarray(u)
            S = lambda u: np.array([0.0])
            self._equation = FVTransportEquation(F, D, S, lambda u: np.diag(u)) if method == "FVM" else FDTransportEquation(F, D, S)
            self._device = compile_model(BURGERS, 2, [nu], method)

    scheme = (sf.FVSchemes if method == "FVM" else sf.FDSchemes)[scheme_name]
    d = sf.Domain.from_size(48, nb[0], nb[1], comps)
    s = sf.Simulation(d, Burgers(), {"a": "sine", "b": "gaussian"}, bcs, scheme)
    rng = np.random.default_rng(3)
    x0 = d.listify_interior(True, [1]) + 0.05 * rng.standard_normal(96)
    P = core.Problem(OracleBurgers(nu), method, scheme_name, np.array(d.positions(interior=True)), nb[0], nb[1], bcs, comps)
    r = s.get_residuals_from_list(x0, True, [1])
    ref = P.residual(x0, True, [1])
    assert np.max(np.abs(r - ref)) <= 1e-12 * np.max(np.abs(ref))
    J = s.jacobian(x0, True, [1]).toarray()
    Jref = P.jacobian_matrix(x0, True, [1]).toarray()
    assert np.max(np.abs(J - Jref)) <= 1e-6 * np.max(np.abs(Jref))
    # and the time stepper runs on it
    s.evolve(0.01)
    assert np.isfinite(d.values()).all()


@pytest.mark.gpu
def test_opaque_model_raises():
    import splitfxm_b200 as sf
    from splitfxm_b200.equations.fd_transport import FDTransportEquation

    class Opaque(sf.Model):
        def __init__(self):
            self._equation = FDTransportEquation(lambda u: u, lambda u: u, lambda u: 0 * u)
            self._device = None

    d = sf.Domain.from_size(20, 1, 1, ["u"])
    s = sf.Simulation(d, Opaque(), {"u": "sine"}, {"u": {"left": "periodic", "right": "periodic"}}, sf.FDSchemes.CENTRAL)
    with pytest.raises(sf.SFXM):
        s.get_residuals_from_list(d.listify_interior(False))
