"""Batched Newton / Split-Newton driver on the device.

The reference delegates to the third-party ``splitnewton`` package (simulation.py:5-6, 515-523), which is not
vendored; the algorithm implemented here is the one specified in ``oracle/newton.py``:
pseudo-transient shift ``J - I/dt`` with SER growth, optional Armijo back-tracking, bound damping, criterion
``|s| / (1e-5 + 1e-6 |x|) < 1``; Split-Newton solves the innermost segment to convergence, then takes one
Newton step on the next segment out.

Every ensemble member carries its own iteration count, criterion and pseudo time step; members that have
converged are masked out of the solve/update kernels, so each member follows exactly the scalar algorithm.
Per iteration the host reads back one flag ("any member still active").
"""
import math

import torch

ABS_TOL = 1e-5
REL_TOL = 1e-6
ARMIJO_C = 1e-4
ARMIJO_MAX_HALVINGS = 10
BOUND_FAC = 0.8


class NewtonDriver:
    def __init__(self, eng, split=False, split_locs=None, dt0=0.0, dtmax=1.0, armijo=False, bounds=None,
                 bound_fac=BOUND_FAC, eps=1e-8, fused=True):
        if dt0 < 0 or dtmax < 0:
            raise ValueError("Must specify positive dt0 and dtmax")
        self.eng = eng
        self.split, self.split_locs = split, split_locs
        _, self.segs = eng.problem(split, split_locs)
        self.dt0, self.dtmax, self.armijo = float(dt0), float(dtmax), bool(armijo)
        self.bound_fac = float(bound_fac)
        self.eps = eps
        dev, B = eng.device, eng.batch
        self.lb = self.ub = None
        if bounds is not None:
            self.lb = torch.as_tensor(bounds[0], dtype=torch.float64, device=dev).reshape(eng.n, -1)
            self.ub = torch.as_tensor(bounds[1], dtype=torch.float64, device=dev).reshape(eng.n, -1)
        self.s = torch.zeros((eng.n, B), dtype=torch.float64, device=dev)
        self.r = torch.empty((eng.n, B), dtype=torch.float64, device=dev)
        self.rt = torch.empty((eng.n, B), dtype=torch.float64, device=dev) if (armijo or dt0 > 0) else None
        self.yt = torch.empty((eng.n, B), dtype=torch.float64, device=dev) if armijo else None
        self.status = torch.zeros(B, dtype=torch.int32, device=dev)
        self.all_active = torch.ones(B, dtype=torch.bool, device=dev)
        # a slab-decomposed mesh (spike.SlabNewtonEngine) reduces the host-side decisions over the ranks
        fused_ok = getattr(eng, "fused_solve_supported", None)
        self.fused = bool(fused_ok(split, split_locs)) if (fused and fused_ok is not None) else False
        self._reduce_min = getattr(eng, "reduce_min", lambda t: t)
        self._reduce_any = getattr(eng, "reduce_any", lambda flag: flag)

    def rows(self, k):
        o, s = self.segs[k]
        return o * self.eng.nx, (o + s) * self.eng.nx

    def check_initial(self, y):
        if self.lb is not None and self._reduce_any(bool(((y < self.lb) | (y > self.ub)).any())):
            raise ValueError("Initial guess is outside the bounds")

    # -- Newton on ONE segment (all segments when unsplit) -------------------------------------------
    def newton_segment(self, y, k, maxiter, parent):
        eng = self.eng
        lo, hi = self.rows(k)
        B, dev = eng.batch, eng.device
        it = torch.zeros(B, dtype=torch.int64, device=dev)
        crit = torch.full((B,), math.inf, dtype=torch.float64, device=dev)
        dt = torch.full((B,), self.dt0, dtype=torch.float64, device=dev)
        while True:
            run = parent & (crit >= 1) & (it < maxiter)
            if not bool(run.any()):
                break
            run32 = run.to(torch.int32)
            shift = (-1.0 / dt) if self.dt0 > 0 else None
            need_rnorm = self.armijo or self.dt0 > 0
            if self.fused:
                # ensembles: the sweeping thread evaluates each cell's block row itself (J never touches HBM)
                eng.newton_solve_fused(y, self.s, self.split, self.split_locs, k, self.eps, shift=shift, active=run32,
                                       status=self.status, r0=self.r if need_rnorm else None)
            else:
                # r = f(y) and the blocks of J = f'(y) that belong to segment k (the only ones this solve reads)
                J = eng.jacobian(y, self.split, self.split_locs, self.eps, r0=self.r, seg_lo=k, seg_hi=k + 1)
                eng.solve(J, self.r, self.s, self.split, self.split_locs, k, k + 1, shift=shift, active=run32,
                          status=self.status)
            rnorm = torch.sqrt(eng.sumsq(self.r, lo, hi)) if need_rnorm else None
            scale = None
            if self.armijo:
                alpha = torch.ones(B, dtype=torch.float64, device=dev)
                done = ~run
                for _ in range(ARMIJO_MAX_HALVINGS):
                    self.yt.copy_(y)
                    eng.axpy(self.yt, self.s, alpha, run32, lo, hi)
                    eng.residual(self.yt, self.rt, self.split, self.split_locs)
                    tn = torch.sqrt(eng.sumsq(self.rt, lo, hi))
                    done = done | (tn <= (1.0 - ARMIJO_C * alpha) * rnorm)
                    if bool(done.all()):
                        break
                    alpha = torch.where(done, alpha, alpha * 0.5)
                scale = alpha
            if self.lb is not None:
                lam = self._bound_lambda(y, lo, hi, scale)
                scale = lam if scale is None else scale * lam
            if scale is not None:
                self.s[lo:hi] = torch.where(run[None, :], self.s[lo:hi] * scale[None, :], self.s[lo:hi])
            ss, yy = eng.update_norms(y, self.s, run32, lo, hi)       # y += s and both norms of the criterion, one pass
            if self.dt0 > 0:
                eng.residual(y, self.rt, self.split, self.split_locs)
                rn = torch.sqrt(eng.sumsq(self.rt, lo, hi))
                grown = torch.where(rn > 0, torch.clamp(dt * rnorm / rn, max=self.dtmax), torch.full_like(dt, self.dtmax))
                dt = torch.where(run, grown, dt)
            c = torch.sqrt(ss) / (ABS_TOL + REL_TOL * torch.sqrt(yy))
            crit = torch.where(run, c, crit)
            it = it + run.to(torch.int64)
        return it

    def _bound_lambda(self, y, lo, hi, scale):
        s = self.s[lo:hi] if scale is None else self.s[lo:hi] * scale[None, :]
        x = y[lo:hi]
        lb, ub = self.lb[lo:hi], self.ub[lo:hi]
        xn = x + s
        one = torch.ones((), dtype=torch.float64, device=y.device)
        up = torch.where(xn > ub, self.bound_fac * (ub - x) / s, one)
        dn = torch.where(xn < lb, self.bound_fac * (lb - x) / s, one)
        return self._reduce_min(torch.minimum(up.min(dim=0).values, dn.min(dim=0).values).clamp(max=1.0))

    # -- Split-Newton over segments k0 .. last --------------------------------------------------------
    def split_newton(self, y, k0, maxiter, parent):
        K = len(self.segs)
        if k0 == K - 1:
            return self.newton_segment(y, k0, maxiter, parent)
        eng = self.eng
        lo, _ = self.rows(k0)
        B, dev = eng.batch, eng.device
        it = torch.zeros(B, dtype=torch.int64, device=dev)
        crit = torch.full((B,), math.inf, dtype=torch.float64, device=dev)
        while True:
            run = parent & (crit >= 1) & (it < maxiter)
            if not bool(run.any()):
                break
            self.split_newton(y, k0 + 1, maxiter, run)      # inner segments to convergence
            self.newton_segment(y, k0, 1, run)              # one step on this segment
            ss, yy = eng.update_norms(y, self.s, None, lo, eng.n, update=False)
            c = torch.sqrt(ss) / (ABS_TOL + REL_TOL * torch.sqrt(yy))
            crit = torch.where(run, c, crit)
            it = it + run.to(torch.int64)
        return it

    def run(self, y, maxiter=math.inf):
        """In place on y [n, batch]; returns per-member iteration counts (int64 [batch])."""
        self.check_initial(y)
        if len(self.segs) == 1:
            return self.newton_segment(y, 0, maxiter, self.all_active)
        return self.split_newton(y, 0, maxiter, self.all_active)
