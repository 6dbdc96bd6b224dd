"""Time stepping on the device for Simulation.evolve (reference: simulation.py:396-472).

* ``euler``   - the reference's hand-rolled loop (simulation.py:455-463), one fused residual+update launch
                per step; reproduces the float-accumulated ``t`` (a trailing tiny step can occur).
* ``rk45``    - scipy.integrate.solve_ivp(method='RK45', t_eval=[t_diff]) restated step for step
                (scipy/integrate/_ivp/rk.py, common.py: select_initial_step, rk_step, error control with
                SAFETY 0.9 / MIN_FACTOR 0.2 / MAX_FACTOR 10, dense output at the end point).  rtol/atol are
                scipy's defaults 1e-3/1e-6 because the reference never forwards its own (simulation.py:468-469).

The state and the 7 stage vectors stay in HBM; per attempted step the host reads back one scalar (the error
norm).  Single-mesh control (batch members advance in lock-step only for Euler).
"""
import ctypes
import math
import os

import numpy as np
import torch

from . import _lib
from .error import SFXM

SAFETY, MIN_FACTOR, MAX_FACTOR = 0.9, 0.2, 10.0


def _tableau():
    from scipy.integrate._ivp.rk import RK45
    return (np.ascontiguousarray(RK45.A, dtype=np.float64), np.ascontiguousarray(RK45.B, dtype=np.float64),
            np.ascontiguousarray(RK45.C, dtype=np.float64), np.ascontiguousarray(RK45.E, dtype=np.float64),
            np.ascontiguousarray(RK45.P, dtype=np.float64))


def _dptr(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_double))


def euler(eng, y, t_diff, max_step, split=False, split_locs=None, on_step=None):
    """Returns (y after t_diff, state of the last RHS evaluation or None, nfev).  simulation.py:455-463."""
    t_current = 0.0
    nfev = 0
    other = torch.empty_like(y)
    while t_current < t_diff:
        delta_t = min(t_diff - t_current, max_step)
        eng.euler_step(y, other, delta_t, split, split_locs)
        y, other = other, y
        nfev += 1
        if on_step is not None:
            on_step(y, t_current)
        t_current += delta_t
    return y, (other if nfev > 0 else None), nfev


class RK45Stepper:
    """rhs(y, out): right-hand side (default: the engine's residual).  reduce(x): combines a local sum of squares
    over all slabs of a decomposed mesh (default: identity); n_global: number of unknowns of the whole mesh."""

    def __init__(self, eng, split=False, split_locs=None, rtol=1e-3, atol=1e-6, max_step=np.inf, rhs=None, reduce=None,
                 n_global=None, on_step=None):
        if eng.batch != 1:
            raise SFXM("adaptive RK45 drives one mesh; ensembles use Euler stepping or per-member Simulations")
        if max_step <= 0:
            raise ValueError("`max_step` must be positive.")
        self.eng, self.split, self.split_locs = eng, split, split_locs
        self.rtol, self.atol, self.max_step = float(rtol), float(atol), float(max_step)
        self.A, self.Bc, self.C, self.E, self.P = _tableau()
        self.n = eng.n
        self.total = eng.n * eng.batch
        dev = eng.device
        self.K = torch.empty((7, eng.n, eng.batch), dtype=torch.float64, device=dev)
        self.ytmp = torch.empty((eng.n, eng.batch), dtype=torch.float64, device=dev)
        self.out1 = torch.empty(eng.batch, dtype=torch.float64, device=dev)
        nbytes = eng.lib.sfxm_scaled_sumsq_workspace_bytes(eng.n, eng.batch)
        self.work = torch.empty(max(nbytes // 8, 1), dtype=torch.float64, device=dev)
        self.nfev = 0
        self.last_rhs_state = None     # state of the last RHS evaluation (the reference's ghosts belong to it)
        self._rhs = rhs
        self._on_step = on_step        # on_step(y, t): called at t = 0 and after every accepted step (scipy's event sites)
        self._reduce = reduce if reduce is not None else (lambda v: v)
        self._n_global = int(n_global) if n_global is not None else eng.n

    def f(self, y, out):
        if self._rhs is not None:
            self._rhs(y, out)
        else:
            self.eng.residual(y, out, self.split, self.split_locs)
        self.nfev += 1

    def _rms(self, a, a2, b):
        """scipy common.norm((a - a2) / (atol + |b| rtol))."""
        eng = self.eng
        _lib.check(eng.lib.sfxm_scaled_diff_sumsq(_lib.ptr(a), _lib.ptr(a2), _lib.ptr(b), self.atol, self.rtol, eng.n,
                                                  eng.batch, _lib.ptr(self.out1), _lib.ptr(self.work), _lib.stream_ptr()))
        _lib.count(2)
        return math.sqrt(self._reduce(float(self.out1[0]))) / self._n_global ** 0.5

    def _lincomb(self, y, coef, nstage, h, out):
        eng = self.eng
        _lib.check(eng.lib.sfxm_rk_lincomb(_lib.ptr(y), _lib.ptr(self.K), self.total, _dptr(coef), nstage, float(h),
                                           _lib.ptr(out), self.total, _lib.stream_ptr()))
        _lib.count()

    def select_initial_step(self, y0, t_bound):
        """common.py select_initial_step (order = 4); uses K[0] = f0 and one more RHS evaluation."""
        interval_length = abs(t_bound)
        if interval_length == 0.0:
            return 0.0
        f0 = self.K[0]
        d0 = self._rms(y0, None, y0)
        d1 = self._rms(f0, None, y0)
        if d0 < 1e-5 or d1 < 1e-5:
            h0 = 1e-6
        else:
            h0 = 0.01 * d0 / d1
        h0 = min(h0, interval_length)
        one = np.array([1.0])
        self._lincomb(y0, one, 1, h0, self.ytmp)            # y1 = y0 + h0 * f0
        self.f(self.ytmp, self.K[1])
        d2 = self._rms(self.K[1], f0, y0) / h0
        if d1 <= 1e-15 and d2 <= 1e-15:
            h1 = max(1e-6, h0 * 1e-3)
        else:
            h1 = (0.01 / max(d1, d2)) ** (1 / 5)
        return min(100 * h0, h1, interval_length, self.max_step)

    # -- step-size control on the device, attempted steps replayed from a CUDA graph -------------------------------
    def device_control_ok(self):
        """The graph path needs the plain engine right-hand side (no slab halo exchange / collective norms inside an
        attempt) and no per-step host callback (bound warnings)."""
        return (self._rhs is None and self._on_step is None and self.eng.device.type == "cuda"
                and os.environ.get("SFXM_RK45_DEVICE", "1")[:1] != "0")

    def _attempt(self):
        """ONE attempted step as a fixed launch sequence (csrc/sfxm_ode.cu): begin, six x (stage combination + RHS),
        error norm, control, commit.  No host decision inside: this is what the CUDA graph holds."""
        eng, lib = self.eng, self.eng.lib
        st = _lib.stream_ptr()
        ctl, flag = _lib.ptr(self._ctl), _lib.ptr(self._flag)
        _lib.check(lib.sfxm_rk45_begin(ctl, flag, st))
        for s in range(1, 6):
            _lib.check(lib.sfxm_rk_lincomb_dev(_lib.ptr(self._y), _lib.ptr(self.K), self.total, _dptr(self.A[s]), s, ctl, flag,
                                               _lib.ptr(self.ytmp), self.total, st))
            eng.residual(self.ytmp, self.K[s], self.split, self.split_locs)
        _lib.check(lib.sfxm_rk_lincomb_dev(_lib.ptr(self._y), _lib.ptr(self.K), self.total, _dptr(self.Bc), 6, ctl, flag,
                                           _lib.ptr(self._ynew), self.total, st))
        eng.residual(self._ynew, self.K[6], self.split, self.split_locs)
        _lib.check(lib.sfxm_rk_error_sumsq_dev(_lib.ptr(self.K), self.total, _dptr(self.E), 7, ctl, _lib.ptr(self._y),
                                               _lib.ptr(self._ynew), self.atol, self.rtol, eng.n, eng.batch,
                                               _lib.ptr(self.out1), _lib.ptr(self.work), st))
        _lib.check(lib.sfxm_rk45_control(ctl, flag, _lib.ptr(self.out1), st))
        _lib.check(lib.sfxm_rk45_commit(_lib.ptr(self._y), _lib.ptr(self._ynew), _lib.ptr(self.K), self.total, _dptr(self.P), 7,
                                        ctl, flag, _lib.ptr(self._result), self.total, st))
        _lib.count(17)

    def integrate_device(self, y, t_bound):
        """integrate() with scipy's controller running on the device: the host computes the initial step (two norms),
        uploads the control block and replays the captured attempt; it reads the flags back once per batch of attempts
        (estimated from the remaining time and the current step), not once per step.  Attempts replayed after the end
        are no-ops on the state.  Same step sequence as integrate() up to the last bits of err ** -0.2."""
        eng = self.eng
        if t_bound <= 0:
            raise ValueError("t_diff must be positive")
        dev = eng.device
        if not hasattr(self, "_y"):
            self._y = torch.empty((eng.n, eng.batch), dtype=torch.float64, device=dev)
            self._ynew = torch.empty_like(self._y)
            self._result = torch.empty_like(self._y)
            self._ctl = torch.zeros(8, dtype=torch.float64, device=dev)
            self._flag = torch.zeros(16, dtype=torch.int32, device=dev)
            self._ctl_host = torch.zeros(8, dtype=torch.float64).pin_memory()
            self._flag_host = torch.zeros(16, dtype=torch.int32).pin_memory()
            self._graph = None
        self._y.copy_(y)
        self.f(self._y, self.K[0])
        h_abs = self.select_initial_step(self._y, t_bound)
        nfev0 = self.nfev
        self._ctl_host[:] = torch.tensor([0.0, h_abs, 0.0, 0.0, float(t_bound), min(self.max_step, 1.7e308), self._n_global ** 0.5,
                                          0.0], dtype=torch.float64)
        self._flag_host.zero_()
        self._flag_host[2] = 1                      # need_begin
        self._ctl.copy_(self._ctl_host, non_blocking=True)
        self._flag.copy_(self._flag_host, non_blocking=True)
        if self._graph is None:
            self._attempt()                         # first attempt eagerly (lazy module loading must not happen in capture)
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                self._attempt()
            self._graph = g
            _lib.LAUNCHES[0] -= 17                  # the captured sequence was recorded, not launched
        while True:
            self._flag_host.copy_(self._flag, non_blocking=True)
            self._ctl_host.copy_(self._ctl, non_blocking=True)
            torch.cuda.current_stream().synchronize()
            if int(self._flag_host[0]):
                break
            t, h = float(self._ctl_host[0]), float(self._ctl_host[1])
            todo = int(min(max(math.ceil((t_bound - t) / h) if h > 0 else 1, 1), 8))      # the step may still grow: few at a time
            for _ in range(todo):
                self._graph.replay()
            _lib.count(17 * todo)
        if int(self._flag_host[5]):
            raise SFXM("Required step size is less than spacing between numbers.")
        self.nfev = nfev0 + int(self._flag_host[6])
        self.last_rhs_state = self._y
        return self._result

    def integrate(self, y, t_bound):
        """Advance y (device [n, 1]) from t = 0 to t_bound; returns the dense-output end value like
        solve_ivp(..., t_eval=[t_bound]).y[:, 0]."""
        if self.device_control_ok():
            return self.integrate_device(y, t_bound)
        eng = self.eng
        if t_bound <= 0:
            raise ValueError("t_diff must be positive")
        y = y.clone()
        y_new = torch.empty_like(y)
        if self._on_step is not None:
            self._on_step(y, 0.0)
        self.f(y, self.K[0])
        h_abs = self.select_initial_step(y, t_bound)
        t = 0.0
        error_exponent = -1 / 5
        result = None
        while t != t_bound:
            min_step = 10 * abs(np.nextafter(t, np.inf) - t)
            if h_abs > self.max_step:
                h_abs = self.max_step
            elif h_abs < min_step:
                h_abs = min_step
            step_accepted = False
            step_rejected = False
            while not step_accepted:
                if h_abs < min_step:
                    raise SFXM("Required step size is less than spacing between numbers.")
                h = h_abs
                t_new = t + h
                if t_new - t_bound > 0:
                    t_new = t_bound
                h = t_new - t
                h_abs = abs(h)
                # rk_step: K[0] holds f(t, y)
                for s in range(1, 6):
                    self._lincomb(y, self.A[s], s, h, self.ytmp)
                    self.f(self.ytmp, self.K[s])
                self._lincomb(y, self.Bc, 6, h, y_new)
                self.f(y_new, self.K[6])
                _lib.check(eng.lib.sfxm_rk_error_sumsq(_lib.ptr(self.K), self.total, _dptr(self.E), 7, float(h), _lib.ptr(y),
                                                       _lib.ptr(y_new), self.atol, self.rtol, eng.n, eng.batch,
                                                       _lib.ptr(self.out1), _lib.ptr(self.work), _lib.stream_ptr()))
                _lib.count(2)
                error_norm = math.sqrt(self._reduce(float(self.out1[0]))) / self._n_global ** 0.5
                if error_norm < 1:
                    if error_norm == 0:
                        factor = MAX_FACTOR
                    else:
                        factor = min(MAX_FACTOR, SAFETY * error_norm ** error_exponent)
                    if step_rejected:
                        factor = min(1, factor)
                    h_abs *= factor
                    step_accepted = True
                else:
                    h_abs *= max(MIN_FACTOR, SAFETY * error_norm ** error_exponent)
                    step_rejected = True
            if t_new == t_bound:
                # t_eval = [t_bound] is served by the dense output of the last step, not by y_new
                result = torch.empty_like(y)
                _lib.check(eng.lib.sfxm_rk_dense_end(_lib.ptr(y), _lib.ptr(self.K), self.total, _dptr(self.P), 7, float(h),
                                                     _lib.ptr(result), self.total, _lib.stream_ptr()))
                _lib.count()
            t = t_new
            y, y_new = y_new, y
            if self._on_step is not None:
                self._on_step(y, t)
            self.last_rhs_state = y
            self.K[0].copy_(self.K[6])      # FSAL: f_new becomes f of the next step
        return result


# ---------------------------------------------------------------------------------------------------------------------
# Implicit BDF: scipy's variable-order step control, the linear algebra on the device
# ---------------------------------------------------------------------------------------------------------------------
class _DeviceJacobian:
    """Stands for the constant Jacobian at y0 (simulation.py:441-443) inside scipy's BDF: `I - c * J` evaluates to a
    token that remembers c, nothing is formed on the host."""
    __array_ufunc__ = None

    def __rmul__(self, c):
        return _ScaledJacobian(float(c))

    __mul__ = __rmul__


class _ScaledJacobian:
    def __init__(self, c):
        self.c = c

    def __rsub__(self, other):
        return _ShiftedSystem(self.c)


class _ShiftedSystem:       # I - c J
    def __init__(self, c):
        self.c = c


class _Identity:
    __array_ufunc__ = None      # numpy scalars defer to __rmul__ instead of broadcasting over the token

    def __sub__(self, other):
        return other.__rsub__(self)

    def __rmul__(self, mu):     # Radau: MU / h * I
        return _ScaledIdentity(mu)

    __mul__ = __rmul__


class _ScaledIdentity:          # mu I
    def __init__(self, mu):
        self.mu = complex(mu) if np.iscomplexobj(mu) else float(mu)

    def __sub__(self, other):
        if not isinstance(other, _DeviceJacobian):
            raise SFXM("device Radau: unexpected linear system")
        return _MuSystem(self.mu)


class _MuSystem:                # mu I - J, mu real or complex
    def __init__(self, mu):
        self.mu = mu


def device_bdf(eng, y0_dev, split, split_locs, eps=1e-8):
    """-> a scipy.integrate.BDF subclass for solve_ivp(method=...) whose Newton systems (I - c J) d = b are solved by the
    block-tridiagonal kernels (sfxm_block_tridiag_solve[_pcr]) on the banded finite-difference Jacobian assembled once at
    y0 on the device.  The reference passes that constant Jacobian to scipy (simulation.py:441-443, 468-469), which
    factorises I - c J with SuperLU on the host; here only the right-hand sides and solutions (n doubles) cross PCIe.
    (I - c J) d = b  <=>  (J - I / c) d = -(b / c): the solver's shifted form with shift = -1 / c, f = b / c.
    In split mode the reference's Jacobian is block-diagonal over the segments (simulation.py:373-392), so the system
    decouples into the per-segment solves the kernel does anyway."""
    from scipy.integrate import BDF
    Jb = eng.jacobian(y0_dev, split, split_locs, eps).clone()
    shift = torch.empty(1, dtype=torch.float64, device=eng.device)
    status = torch.zeros(1, dtype=torch.int32, device=eng.device)
    sol = torch.zeros((eng.n, 1), dtype=torch.float64, device=eng.device)

    def lu(A):
        if not isinstance(A, _ShiftedSystem):
            raise SFXM("device BDF: unexpected linear system")
        return A

    def solve_lu(LU, b):
        f = eng.to_device(np.ascontiguousarray(b, dtype=np.float64) / LU.c)
        shift.fill_(-1.0 / LU.c)
        eng.solve(Jb, f, sol, split, split_locs, shift=shift, status=status)
        return sol.cpu().numpy().reshape(-1)

    class DeviceBDF(BDF):
        def __init__(self, fun, t0, y0, t_bound, **kw):
            super().__init__(fun, t0, y0, t_bound, **kw)
            self.J = _DeviceJacobian()
            self.I = _Identity()
            self.lu = lu
            self.solve_lu = solve_lu
            self.LU = None

    return DeviceBDF


# ---------------------------------------------------------------------------------------------------------------------
# Implicit Radau IIA: scipy's step control, the real and the COMPLEX shifted systems on the device
# ---------------------------------------------------------------------------------------------------------------------
def device_radau(eng, y0_dev, split, split_locs, eps=1e-8):
    """-> a scipy.integrate.Radau subclass for solve_ivp(method=...), or None when a segment is too wide (below).

    scipy's Radau (radau.py `_step_impl`, `solve_collocation_system`) factorises two matrices per step size with the
    constant Jacobian at y0 the reference hands it (simulation.py:441-443, 468-469): MU_REAL / h I - J (real) and
    MU_COMPLEX / h I - J (complex), and solves with them in every simplified-Newton iteration.  Here J stays in the banded
    device layout and both are solved by the block-tridiagonal kernels:
      * real:    (mu I - J) x = b        <=>  (J - mu I) x = -b                    (the solver's shifted form, shift = -mu);
      * complex: mu = a + i b, z = x + i y, rhs = p + i q.  Interleaving (x, y) per cell turns every s x s block B of a
        segment into the 2s x 2s real block [[B, 0], [0, B]] and the diagonal block into [[B_d, b I], [-b I, B_d]] with
        shift -a: still block-tridiagonal (cyclic wrap blocks included), solved by the same kernel with 2s unknowns per cell.
    The kernels hold blocks up to 4 x 4, so segments of 1 or 2 variables are served (any split run of <= 2-variable
    segments, unsplit runs of <= 2 variables); wider segments keep scipy's host factorisation (the caller falls back).
    Only right-hand sides and solutions cross PCIe."""
    import ctypes
    from scipy.integrate import Radau
    from . import _lib
    P, segs = eng.problem(split, split_locs)
    if eng.W != 3 or eng.batch != 1 or any(2 * s > _lib.MAX_NV for _, s in segs):
        return None
    dev, nx = eng.device, eng.nx
    Jb = eng.jacobian(y0_dev, split, split_locs, eps).clone()          # [nx, 3, nv, nv, 1]
    shift = torch.empty(1, dtype=torch.float64, device=dev)
    status = torch.zeros(1, dtype=torch.int32, device=dev)
    sol = torch.zeros((eng.n, 1), dtype=torch.float64, device=dev)
    lib = eng.lib

    class _ComplexSegment:
        """One segment's 2s-unknowns-per-cell real form: descriptor, J2 / rhs / solution buffers, workspace."""

        def __init__(self, o, s):
            self.o, self.s = o, s
            P2 = _lib.Problem.from_buffer_copy(P)
            P2.nv = 2 * s
            for j in range(2 * s):
                src = o + (j % s)
                P2.var_base[j], P2.var_stride[j], P2.var_off[j], P2.var_seg[j] = 0, 2 * s, j, 0
                P2.bc_ltype[j], P2.bc_rtype[j] = P.bc_ltype[src], P.bc_rtype[src]
            self.P2 = P2
            self.J2 = torch.zeros((nx, 3, 2 * s, 2 * s, 1), dtype=torch.float64, device=dev)
            B = Jb[:, :, o:o + s, o:o + s, 0]
            self.J2[:, :, :s, :s, 0] = B
            self.J2[:, :, s:, s:, 0] = B
            self.f2 = torch.empty((nx * 2 * s, 1), dtype=torch.float64, device=dev)
            self.s2 = torch.zeros((nx * 2 * s, 1), dtype=torch.float64, device=dev)
            self.work = torch.empty(max(lib.sfxm_pcr_workspace_bytes(ctypes.byref(P2)) // 8, 1), dtype=torch.float64, device=dev)
            self.beta = 0.0

        def solve(self, mu, rhs_seg):
            """rhs_seg: complex [nx, s] device tensor -> complex [nx, s] with (mu I - J_seg) z = rhs."""
            s = self.s
            if mu.imag != self.beta:                              # the coupling entries of the diagonal blocks
                r = torch.arange(s, device=dev)
                self.J2[:, 1, r, s + r, 0] += mu.imag - self.beta
                self.J2[:, 1, s + r, r, 0] -= mu.imag - self.beta
                self.beta = mu.imag
            # (J2 + shift I) v = -f2 with f2 = (Re rhs, Im rhs) per cell, shift = -Re mu
            self.f2.copy_(torch.cat([rhs_seg.real, rhs_seg.imag], dim=1).reshape(-1, 1))
            shift.fill_(-mu.real)
            _lib.check(lib.sfxm_block_tridiag_solve_pcr(ctypes.byref(self.P2), _lib.ptr(self.J2), _lib.ptr(self.f2), _lib.ptr(shift),
                                                        None, 0, 1, _lib.ptr(self.s2), _lib.ptr(self.work), _lib.ptr(status),
                                                        _lib.stream_ptr()))
            _lib.count()
            v = self.s2.reshape(nx, 2 * s)
            return torch.complex(v[:, :s], v[:, s:])

    csegs = [_ComplexSegment(o, s) for (o, s) in segs]

    def solve_lu(LU, b):
        if isinstance(LU.mu, complex):
            bd = torch.from_numpy(np.ascontiguousarray(b, dtype=np.complex128)).to(dev)
            out = torch.empty(eng.n, dtype=torch.complex128, device=dev)
            for c in csegs:       # segment k occupies rows [o nx, (o + s) nx) of the flat vector, cell-major inside
                lo, hi = c.o * nx, (c.o + c.s) * nx
                out[lo:hi] = c.solve(LU.mu, bd[lo:hi].reshape(nx, c.s)).reshape(-1)
            return out.cpu().numpy()
        f = eng.to_device(np.ascontiguousarray(b, dtype=np.float64))
        shift.fill_(-LU.mu)
        eng.solve(Jb, f, sol, split, split_locs, shift=shift, status=status)
        return sol.cpu().numpy().reshape(-1)

    class DeviceRadau(Radau):
        def __init__(self, fun, t0, y0, t_bound, **kw):
            super().__init__(fun, t0, y0, t_bound, **kw)
            self.J = _DeviceJacobian()
            self.I = _Identity()

            def lu(A):                      # nothing to factorise: the shifted solve takes mu itself
                if not isinstance(A, _MuSystem):
                    raise SFXM("device Radau: unexpected linear system")
                self.nlu += 1
                return A
            self.lu = lu
            self.solve_lu = solve_lu
            self.LU_real = self.LU_complex = None

    return DeviceRadau
