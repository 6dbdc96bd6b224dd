"""The nonlinear steady-state cases really are nonlinear: the specification (oracle/newton.py) needs several iterations
on them and takes its Armijo / bound / pseudo-transient branches - asserted from its trace, on CPU.  The CUDA path is
compared with these runs in tests/test_gpu_newton_nonlinear.py, the device driver's control logic in
tests/test_newton_driver_cpu.py."""
import numpy as np
import pytest

from nonlinear_cases import CASES, oracle_case, oracle_solve


@pytest.mark.parametrize("name", sorted(CASES))
def test_case_exercises_the_nonlinear_branches(name):
    P, U0, c = oracle_case(name)
    x0, xf, it, trace = oracle_solve(name)
    e = c["expect"]
    assert it >= e["min_it"] and it < 200
    assert np.max(np.abs(P.residual(xf, c["split"], c["locs"]))) < 1e-8          # converged to a steady state
    if e.get("halvings"):
        assert any(t["alpha"] < 1.0 for t in trace), "Armijo never back-tracked"
    if e.get("clips"):
        assert any(t["lam"] < 1.0 for t in trace), "no step was clipped at the bounds"
    if e.get("ptc"):
        dts = [t["dt"] for t in trace]
        assert dts[0] == c["kw"]["dt0"] and max(dts) > 10 * dts[0], "the pseudo time step never grew"
    # the count is not on a knife edge: the last criterion is well below 1, the one before well above
    assert trace[-1]["crit"] < 0.7


@pytest.mark.parametrize("name", sorted(CASES))
def test_case_is_not_rounding_sensitive(name):
    """The finite-difference Jacobian's round-off noise is a discontinuous function of the iterate's last bits, so two
    correct implementations whose linear solves round differently can drift apart on wandering damped iterations
    (Burgers nu = 0.03: 25 vs 22 iterations between SuperLU and LAPACK solves of this very specification).  The parity
    cases are chosen where that does not happen: both roundings give the same count and the same iterate to 1e-11."""
    _, xa, ita, _ = oracle_solve(name)
    _, xb, itb, _ = oracle_solve(name, dense=True)
    assert ita == itb
    assert np.max(np.abs(xa - xb)) <= 1e-11 * np.max(np.abs(xa))
