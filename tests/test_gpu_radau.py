"""Device Radau (splitfxm_b200/ode.py device_radau): the real and the complex shifted systems of scipy's Radau solved by the
block-tridiagonal kernels on the banded device Jacobian.  The trajectory parity against the reference's fixtures is in
test_gpu_parity.py::test_evolve_matches_reference (evolve_readme_radau_split, evolve_radau_split1_fd); here: which runs are
served on the device, and both solves against dense LAPACK solves of the same matrices."""
import numpy as np
import pytest

from helpers import load, product_simulation

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", ["evolve_readme_radau_split", "evolve_radau_split1_fd"])
def test_shifted_solves_match_dense_solves(name):
    from scipy.sparse import identity
    from splitfxm_b200.ode import _MuSystem, device_radau
    z = load(name)
    s, d, m = product_simulation(z, z["x_init"], z["U_init"])
    eng = s._engine()
    y = s._state(eng, m["split"], m["locs"])
    cls = device_radau(eng, y, m["split"], m["locs"])
    assert cls is not None, "segments of one or two variables must be served on the device"
    n = eng.n
    y0 = y.cpu().numpy().reshape(-1)
    J = s.jacobian(y0, m["split"], m["locs"]).toarray()
    solver = cls(lambda t, v: np.zeros_like(v), 0.0, y0, 1.0, jac=identity(n, format="csc"))
    rng = np.random.default_rng(5)
    b = rng.standard_normal(n)
    for mu in (3.7, 41.0):
        want = np.linalg.solve(mu * np.eye(n) - J, b)
        got = solver.solve_lu(_MuSystem(mu), b)
        assert np.max(np.abs(got - want)) <= 1e-12 * np.max(np.abs(want))
    bc = b + 1j * rng.standard_normal(n)
    for mu in (2.6 + 3.1j, 26.0 - 30.5j, 2.6 + 3.1j):          # back to the first: the coupling entries are re-set
        want = np.linalg.solve(mu * np.eye(n) - J, bc)
        got = solver.solve_lu(_MuSystem(mu), bc)
        assert np.max(np.abs(got - want)) <= 1e-12 * np.max(np.abs(want))


def test_wide_segments_fall_back_to_the_host_factorisation():
    from splitfxm_b200.ode import device_radau
    z = load("evolve_readme_radau")                              # unsplit, three variables: 6 x 6 real blocks
    s, d, m = product_simulation(z, z["x_init"], z["U_init"])
    eng = s._engine()
    assert device_radau(eng, s._state(eng, False, None), False, None) is None
    s.evolve(float(z["t"]), False, None, method="Radau")
    assert np.max(np.abs(np.array(d.values()) - z["U_all"])) <= 1e-13 * np.max(np.abs(z["U_all"]))
