"""The reference's kernel-level functions (fluxes, diffusion_fluxes, Dx, D2x, dx, d2x) with its own known answers
(tests/test_flux.py:54-168, tests/test_derivative.py:17-91), evaluated by csrc/sfxm_stencil_api.cu."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _cells(vals, xs=None):
    import splitfxm_b200 as sf
    xs = xs if xs is not None else range(len(vals))
    return [sf.Cell(float(x), np.array(v, dtype=np.float64)) for x, v in zip(xs, vals)]


def test_finite_volume_fluxes():
    import splitfxm_b200 as sf
    from splitfxm_b200.flux import diffusion_fluxes, fluxes
    F = lambda u: 2.0 * u
    dfdu = lambda u: np.diag([2.0] * len(u))
    c3 = _cells([[1.0, 2.0], [3.0, 4.0], [2.0, 6.0]], [0.0, 1.0, 2.0])
    ul, uc, ur = (c.values() for c in c3)
    Fw, Fe = fluxes(F, c3, sf.FVSchemes.LAX_FRIEDRICHS, dfdu)
    assert np.allclose(Fw, 0.5 * (F(ul) + F(uc)) - 0.5 * 2 * (uc - ul))
    assert np.allclose(Fe, 0.5 * (F(uc) + F(ur)) - 0.5 * 2 * (ur - uc))
    Fw, Fe = fluxes(F, c3, sf.FVSchemes.CENTRAL)
    assert np.allclose(Fw, 0.5 * (F(ul) + F(uc))) and np.allclose(Fe, 0.5 * (F(uc) + F(ur)))
    Fw, Fe = fluxes(F, c3, sf.FVSchemes.UPWIND, dfdu)                    # positive speeds: upwind is the left state
    assert np.allclose(Fw, F(ul)) and np.allclose(Fe, F(uc))
    Fn = lambda u: -2.0 * u
    Fw, Fe = fluxes(Fn, c3, sf.FVSchemes.UPWIND, lambda u: np.diag([-2.0] * len(u)))
    assert np.allclose(Fw, Fn(uc)) and np.allclose(Fe, Fn(ur))
    D = lambda u: 0.1 * u
    Dw, De = diffusion_fluxes(D, c3, sf.FVSchemes.LAX_FRIEDRICHS)
    assert np.allclose(Dw, (D(uc) - D(ul)) / (0.5 * 1.0)) and np.allclose(De, (D(ur) - D(uc)) / (0.5 * 1.0))
    for bad in (lambda: fluxes(F, c3, sf.FVSchemes.LAX_FRIEDRICHS), lambda: fluxes(F, c3, sf.FVSchemes.UPWIND),
                lambda: fluxes(F, c3[:1], sf.FVSchemes.LAX_FRIEDRICHS, dfdu), lambda: fluxes(F, c3, "UNSUPPORTED", dfdu),
                lambda: diffusion_fluxes(D, c3[:1], sf.FVSchemes.LAX_FRIEDRICHS), lambda: diffusion_fluxes(D, c3, "UNSUPPORTED")):
        with pytest.raises(sf.SFXM):
            bad()


def test_finite_difference_derivatives_of_linear_data():
    import splitfxm_b200 as sf
    from splitfxm_b200.derivatives import D2x, Dx, d2x, dx
    c3 = _cells([[1.0], [2.0], [3.0]])
    c4 = _cells([[1.0], [2.0], [3.0], [4.0]])
    ident = lambda v: v
    assert np.allclose(Dx(ident, c3, sf.FDSchemes.CENTRAL), [1.0])
    assert np.allclose(Dx(ident, c4, sf.FDSchemes.RIGHT_BIAS), [1.0])
    assert D2x(ident, c3, sf.FDSchemes.CENTRAL) == 0.0 and D2x(ident, c4, sf.FDSchemes.RIGHT_BIAS) == 0.0
    assert dx(np.array([1.0, 2, 3]), c3, sf.FDSchemes.CENTRAL) == 1.0
    assert dx(np.array([1.0, 2, 3, 4]), c4, sf.FDSchemes.RIGHT_BIAS) == 1.0
    assert d2x(np.array([1.0, 2, 3]), c3, sf.FDSchemes.CENTRAL) == 0.0
    assert d2x(np.array([1.0, 2, 3, 4]), c4, sf.FDSchemes.RIGHT_BIAS) == 0.0
    quad = _cells([[0.0, 1.0], [1.0, 2.0], [4.0, 5.0]], [0.0, 1.0, 2.0])   # x^2 and x^2 + 1 on a unit mesh
    assert np.allclose(D2x(ident, quad, sf.FDSchemes.CENTRAL), [2.0, 2.0])
    with pytest.raises(sf.SFXM):
        Dx(ident, c3[:2], sf.FDSchemes.CENTRAL)
