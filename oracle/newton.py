"""Newton / Split-Newton CPU specification - TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

PARITY UNPINNED.  The reference calls the third-party package ``splitnewton`` (setup.py:45,
requirements.txt:5 - no version pin) at simulation.py:5-6, 515-523:

    newton(_f, _jac, x0, sparse=, dt0=, dtmax=, armijo=, bounds=)              -> (x, _, iter)
    split_newton(_f, _jac, x0, locs, sparse=, dt0=, dtmax=, armijo=, bounds=)  -> (x, _, iter)

That package is not vendored in /root/reference and cannot be installed here, and the reference's own tests
only assert ``iter >= 0`` (tests/test_simulation.py:112-125).  This file therefore DEFINES the behaviour both
the CPU baseline and the CUDA path implement: it follows the algorithm text in the
reference README (README.md:13-35) - innermost segment Newton-iterated to convergence with the outer segments
frozen, then one Newton step on the next segment out, repeated until the whole-system criterion is met - and
the call signature above.  Iteration-count parity claims in this repository are against THIS restatement.
"""
import numpy as np
from scipy.sparse import csc_matrix, identity, issparse
from scipy.sparse.linalg import spsolve

ABS_TOL = 1e-5
REL_TOL = 1e-6
ARMIJO_C = 1e-4
ARMIJO_MAX_HALVINGS = 10
BOUND_FAC = 0.8


def criterion(x, s):
    return np.linalg.norm(s) / (ABS_TOL + REL_TOL * np.linalg.norm(x))


def _solve(A, b, sparse):
    if issparse(A):
        return spsolve(csc_matrix(A), b)
    return np.linalg.solve(np.asarray(A), b)


def bound_step(x, s, bounds, bound_fac=BOUND_FAC):
    """Largest lam <= 1 keeping lb <= x + lam s <= ub; a clipped step is shortened by bound_fac."""
    if bounds is None:
        return 1.0
    lb, ub = np.asarray(bounds[0], dtype=float), np.asarray(bounds[1], dtype=float)
    lam = 1.0
    xn = x + s
    over = xn > ub
    if np.any(over):
        lam = min(lam, np.min(bound_fac * (ub[over] - x[over]) / s[over]))
    under = xn < lb
    if np.any(under):
        lam = min(lam, np.min(bound_fac * (lb[under] - x[under]) / s[under]))
    return lam


def newton(f, J, x0, maxiter=np.inf, sparse=False, dt0=0.0, dtmax=1.0, armijo=False, bounds=None,
           bound_fac=BOUND_FAC, trace=None):
    """trace (tests only): a list that receives one dict per iteration - n (unknowns of this solve), alpha (Armijo
    factor), lam (bound damping), dt (pseudo time step used), crit."""
    if dt0 < 0 or dtmax < 0:
        raise ValueError("Must specify positive dt0 and dtmax")
    x = np.array(x0, dtype=np.float64)
    if bounds is not None:
        if np.any(x < np.asarray(bounds[0])) or np.any(x > np.asarray(bounds[1])):
            raise ValueError("Initial guess is outside the bounds")
    n = x.shape[0]
    dt = dt0
    it = 0
    crit = np.inf
    s = np.zeros_like(x)
    while crit >= 1 and it < maxiter:
        r = np.asarray(f(x), dtype=np.float64)
        A = J(x)
        dt_used, alpha = dt, 1.0
        if dt0 > 0:
            A = A - (1.0 / dt) * (identity(n, format="csr") if issparse(A) else np.eye(n))
        s = _solve(A, -r, sparse)
        rnorm = np.linalg.norm(r)
        if armijo:
            for _ in range(ARMIJO_MAX_HALVINGS):
                if np.linalg.norm(f(x + alpha * s)) <= (1.0 - ARMIJO_C * alpha) * rnorm:
                    break
                alpha *= 0.5
            s = alpha * s
        lam = bound_step(x, s, bounds, bound_fac)
        if lam != 1.0:
            s = lam * s
        x = x + s
        if dt0 > 0:
            rn = np.linalg.norm(f(x))
            dt = min(dt * rnorm / rn, dtmax) if rn > 0 else dtmax
        crit = criterion(x, s)
        it += 1
        if trace is not None:
            trace.append(dict(n=n, alpha=alpha, lam=float(lam), dt=dt_used, crit=float(crit)))
    return x, s, it


def split_newton(f, J, x0, locs, maxiter=np.inf, sparse=False, dt0=0.0, dtmax=1.0, armijo=False, bounds=None,
                 bound_fac=BOUND_FAC, trace=None):
    """locs ascending, in flat-vector units (simulation.py:520: num_points * split_loc)."""
    x0 = np.array(x0, dtype=np.float64)
    if len(locs) == 0:
        return newton(f, J, x0, maxiter, sparse, dt0, dtmax, armijo, bounds, bound_fac, trace)
    loc = int(locs[0])
    xa, xb = x0[:loc].copy(), x0[loc:].copy()
    ba = bb = None
    if bounds is not None:
        lb, ub = np.asarray(bounds[0], dtype=float), np.asarray(bounds[1], dtype=float)
        ba, bb = [lb[:loc], ub[:loc]], [lb[loc:], ub[loc:]]
    it = 0
    crit = np.inf
    sa, sb = np.zeros_like(xa), np.zeros_like(xb)
    rest = [int(l) - loc for l in locs[1:]]
    while crit >= 1 and it < maxiter:
        # inner (right) part to convergence with the outer (left) part frozen
        def fb(y): return np.asarray(f(np.concatenate([xa, y])))[loc:]
        def Jb(y): return J(np.concatenate([xa, y]))[loc:, loc:]
        if rest:
            xb, sb, _ = split_newton(fb, Jb, xb, rest, maxiter, sparse, dt0, dtmax, armijo, bb, bound_fac, trace)
        else:
            xb, sb, _ = newton(fb, Jb, xb, maxiter, sparse, dt0, dtmax, armijo, bb, bound_fac, trace)
        # one Newton step on the outer part with the inner part frozen
        def fa(y): return np.asarray(f(np.concatenate([y, xb])))[:loc]
        def Ja(y): return J(np.concatenate([y, xb]))[:loc, :loc]
        xa, sa, _ = newton(fa, Ja, xa, 1, sparse, dt0, dtmax, armijo, ba, bound_fac, trace)
        x = np.concatenate([xa, xb])
        s = np.concatenate([sa, sb])
        crit = criterion(x, s)
        it += 1
    return np.concatenate([xa, xb]), np.concatenate([sa, sb]), it
