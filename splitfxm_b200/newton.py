"""Batched Newton / Split-Newton driver on the device.

The reference delegates to the third-party ``splitnewton`` package (simulation.py:5-6, 515-523), which is not
vendored; the algorithm implemented here is the one specified in ``oracle/newton.py``:
pseudo-transient shift ``J - I/dt`` with SER growth, optional Armijo back-tracking, bound damping, criterion
``|s| / (1e-5 + 1e-6 |x|) < 1``; Split-Newton solves the innermost segment to convergence, then takes one
Newton step on the next segment out.

Every ensemble member carries its own iteration count, criterion and pseudo time step; members that have
converged are masked out of the solve/update kernels, so each member follows exactly the scalar algorithm.

Per iteration the per-member bookkeeping is two small kernels (csrc/sfxm_newton_ctl.cu: ``_begin`` - who still runs,
the pseudo-transient shift, one "any member active" flag for the host; ``_end`` - criterion, counts, time-step growth);
on CPU tensors (the driver's control logic is tested on CPU over an oracle-backed engine) the same bookkeeping runs as
tensor expressions.

The host never stalls the device on that flag: it is copied to pinned memory behind an event, the iteration's kernels -
all masked by the per-member run flags, so an iteration nobody needs is a no-op - are launched right away, and only then
does the host look at the flag (``_peek`` before, ``_wait`` after).  When the device is already idle the flag is known
at once and nothing is launched speculatively.
"""
import ctypes
import math
import os

import torch

ABS_TOL = 1e-5
REL_TOL = 1e-6
ARMIJO_C = 1e-4
ARMIJO_MAX_HALVINGS = 10
BOUND_FAC = 0.8


class NewtonDriver:
    def __init__(self, eng, split=False, split_locs=None, dt0=0.0, dtmax=1.0, armijo=False, bounds=None,
                 bound_fac=BOUND_FAC, eps=1e-8, fused=True, native=True):
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
        # per segment: sum s^2 of the member's last step there and sum y^2 of its current state there - an enclosing
        # Split-Newton level forms its criterion from these instead of another pass over (y, s)
        self.seg_ss = torch.zeros((len(self.segs), B), dtype=torch.float64, device=dev)
        self.seg_yy = torch.zeros((len(self.segs), B), dtype=torch.float64, device=dev)
        # a slab-decomposed mesh (spike.SlabNewtonEngine) reduces the host-side decisions over the ranks
        fused_ok = getattr(eng, "fused_solve_supported", None)
        self.fused = bool(fused_ok(split, split_locs)) if (fused and fused_ok is not None) else False
        self._reduce_min = getattr(eng, "reduce_min", lambda t: t)
        self._reduce_any = getattr(eng, "reduce_any", lambda flag: flag)
        # per-member bookkeeping on the device through the C ABI when the engine has one (CUDA), else tensor expressions
        self._lib = getattr(eng, "lib", None) if dev.type == "cuda" else None
        # on a slab-decomposed mesh every iteration contains collectives: all ranks must take the same decision, so the flag
        # is always awaited there (whether an event has completed yet differs from rank to rank)
        self._speculate = not hasattr(eng, "comm")
        if self._lib is not None:
            self._any = torch.zeros(1, dtype=torch.int32, device=dev)
            # one pinned flag word + one event per loop level: an enclosing level's ticket stays valid while the loops
            # inside it issue theirs
            self._any_host = torch.zeros(4 * (len(self.segs) + 1), dtype=torch.int32).pin_memory()
            self._states, self._reset, self._slots = {}, False, {}
            self._shift = torch.empty(B, dtype=torch.float64, device=dev) if dt0 > 0 else None
        # plain runs (full steps, no pseudo-transient term / Armijo / bounds) on one process: the whole loop nest is driven
        # from inside the library (sfxm_newton_run, csrc/sfxm_newton_driver.cu) - the same calls in the same order, without
        # the interpreter between the launches
        self.native = bool(native and self._lib is not None and self._speculate and self.dt0 == 0 and not self.armijo
                           and self.lb is None and os.environ.get("SFXM_NEWTON_NATIVE", "1")[:1] != "0")
        self._plan = None
        self._plan_cache = None

    def rows(self, k):
        o, s = self.segs[k]
        return o * self.eng.nx, (o + s) * self.eng.nx

    def check_initial(self, y):
        if self.lb is not None and self._reduce_any(bool(((y < self.lb) | (y > self.ub)).any())):
            raise ValueError("Initial guess is outside the bounds")

    # -- per-member bookkeeping --------------------------------------------------------------------------
    def _new_state(self, with_dt, key):
        """(it, crit, dt, run) of one loop level, all [batch].  On the device the buffers of a level are allocated once
        and reset by the loop's first sfxm_newton_begin; a level's previous counts have been consumed by then (the
        caller reads them only after the whole run)."""
        B, dev = self.eng.batch, self.eng.device
        if self._lib is not None:
            st = self._states.get(key)
            if st is None:
                st = (torch.zeros(B, dtype=torch.int64, device=dev), torch.empty(B, dtype=torch.float64, device=dev),
                      torch.empty(B, dtype=torch.float64, device=dev) if with_dt else None,
                      torch.zeros(B, dtype=torch.int32, device=dev))
                self._states[key] = st
            self._reset = True
            return st
        it = torch.zeros(B, dtype=torch.int64, device=dev)
        crit = torch.full((B,), math.inf, dtype=torch.float64, device=dev)
        dt = torch.full((B,), self.dt0, dtype=torch.float64, device=dev) if with_dt else None
        run = torch.zeros(B, dtype=torch.int32, device=dev)
        return it, crit, dt, run

    def _begin(self, parent, it, crit, dt, run, maxiter, key=None):
        """run <- parent & (crit >= 1) & (it < maxiter) (int32 mask, in place); -> (ticket for _peek / _wait, shift or
        None).  On the device the "any member runs" flag travels to the loop level's pinned word behind its event."""
        use_shift = dt is not None and self.dt0 > 0
        if self._lib is not None:
            from . import _lib
            shift = self._shift if use_shift else None
            reset, self._reset = self._reset, False
            _lib.check(self._lib.sfxm_newton_begin(self.eng.batch, _lib.ptr(parent), _lib.ptr(crit), _lib.ptr(it),
                                                   float(min(maxiter, 1e300)), _lib.ptr(dt) if use_shift else None, self.dt0,
                                                   1 if reset else 0, _lib.ptr(run), _lib.ptr(shift), _lib.ptr(self._any),
                                                   _lib.stream_ptr()))
            _lib.count()
            slot = self._slots.get(key)
            if slot is None:
                slot = self._slots[key] = (len(self._slots), torch.cuda.Event())
            self._any_host[slot[0]:slot[0] + 1].copy_(self._any, non_blocking=True)
            slot[1].record()
            return slot, shift
        r = (crit >= 1) & (it < maxiter)
        if parent is not None:
            r = r & (parent != 0)
        run.copy_(r.to(torch.int32))
        return bool(r.any()), ((-1.0 / dt) if use_shift else None)

    def _peek(self, ticket):
        """True / False if it is already known whether any member runs, None while the flag is still in flight."""
        if isinstance(ticket, bool):
            return ticket
        if not self._speculate:
            return self._wait(ticket)
        return bool(int(self._any_host[ticket[0]])) if ticket[1].query() else None

    def _wait(self, ticket):
        if isinstance(ticket, bool):
            return ticket
        ticket[1].synchronize()
        return bool(int(self._any_host[ticket[0]]))

    def _end(self, run, ss, yy, it, crit, dt=None, r_old_sq=None, r_new_sq=None, nsum=1, keep=None):
        """crit, it (and the pseudo time step) of the members that ran, in place.  ss, yy: [batch], or [nsum, batch] whose
        rows are summed (the kept sums of the segments of an enclosing Split-Newton level); keep = (ss_row, yy_row)
        receives the sums of the members that ran."""
        if self._lib is not None:
            from . import _lib
            _lib.check(self._lib.sfxm_newton_end(self.eng.batch, _lib.ptr(run), _lib.ptr(ss), _lib.ptr(yy), int(nsum), ABS_TOL,
                                                 REL_TOL, _lib.ptr(crit), _lib.ptr(it), _lib.ptr(dt), _lib.ptr(r_old_sq),
                                                 _lib.ptr(r_new_sq), self.dtmax, _lib.ptr(keep[0]) if keep else None,
                                                 _lib.ptr(keep[1]) if keep else None, _lib.stream_ptr()))
            _lib.count()
            return
        on = run != 0
        if nsum > 1:
            a, b = ss[0].clone(), yy[0].clone()
            for q in range(1, nsum):
                a, b = a + ss[q], b + yy[q]
            ss, yy = a, b
        if keep is not None:
            keep[0].copy_(torch.where(on, ss, keep[0]))
            keep[1].copy_(torch.where(on, yy, keep[1]))
        c = torch.sqrt(ss) / (ABS_TOL + REL_TOL * torch.sqrt(yy))
        crit.copy_(torch.where(on, c, crit))
        it.add_(on.to(torch.int64))
        if dt is not None:
            rn, rnorm = torch.sqrt(r_new_sq), torch.sqrt(r_old_sq)
            grown = torch.where(rn > 0, torch.clamp(dt * rnorm / rn, max=self.dtmax), torch.full_like(dt, self.dtmax))
            dt.copy_(torch.where(on, grown, dt))

    # -- Newton on ONE segment (all segments when unsplit) -------------------------------------------
    def newton_segment(self, y, k, maxiter, parent):
        eng = self.eng
        lo, hi = self.rows(k)
        B, dev = eng.batch, eng.device
        ptc = self.dt0 > 0
        level = ("seg", k, maxiter == 1)
        it, crit, dt, run32 = self._new_state(ptc, level)
        need_rnorm = self.armijo or ptc
        while True:
            ticket, shift = self._begin(parent, it, crit, dt, run32, maxiter, level)
            known = self._peek(ticket)
            if known is False:
                break
            # everything below is masked by run32: launched before the flag is known, it is a no-op if nobody runs
            stepped = None
            if self.fused and not self.armijo and self.lb is None:
                # ensembles, full steps: the sweeping thread evaluates each cell's block row itself (J never touches HBM),
                # applies the step and accumulates the criterion's sums on its way back
                stepped = eng.newton_step_fused(y, self.s, self.split, self.split_locs, k, self.eps, shift=shift,
                                                active=run32, status=self.status, r0=self.r if need_rnorm else None)
            elif self.fused:
                eng.newton_solve_fused(y, self.s, self.split, self.split_locs, k, self.eps, shift=shift, active=run32,
                                       status=self.status, r0=self.r if need_rnorm else None)
            else:
                # r = f(y) and the blocks of J = f'(y) that belong to segment k (the only ones this solve reads)
                J = eng.jacobian(y, self.split, self.split_locs, self.eps, r0=self.r, seg_lo=k, seg_hi=k + 1)
                eng.solve(J, self.r, self.s, self.split, self.split_locs, k, k + 1, shift=shift, active=run32,
                          status=self.status)
            r_old_sq = eng.sumsq(self.r, lo, hi) if need_rnorm else None
            scale = None
            if self.armijo:
                run = run32 != 0
                rnorm = torch.sqrt(r_old_sq)
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
                run = run32 != 0
                self.s[lo:hi] = torch.where(run[None, :], self.s[lo:hi] * scale[None, :], self.s[lo:hi])
            if stepped is not None:
                ss, yy = stepped
            else:
                ss, yy = eng.update_norms(y, self.s, run32, lo, hi)   # y += s and both norms of the criterion, one pass
            r_new_sq = None
            if ptc:
                eng.residual(y, self.rt, self.split, self.split_locs)
                r_new_sq = eng.sumsq(self.rt, lo, hi)
            self._end(run32, ss, yy, it, crit, dt, r_old_sq if ptc else None, r_new_sq, keep=(self.seg_ss[k], self.seg_yy[k]))
            if maxiter == 1:          # one step (the outer segments of Split-Newton): nobody can run a second time
                break
            if known is None and not self._wait(ticket):
                break
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
        level = ("split", k0)
        it, crit, _, run32 = self._new_state(False, level)
        while True:
            ticket, _ = self._begin(parent, it, crit, None, run32, maxiter, level)
            known = self._peek(ticket)
            if known is False:
                break
            self.split_newton(y, k0 + 1, maxiter, run32)      # inner segments to convergence
            self.newton_segment(y, k0, 1, run32)              # one step on this segment
            # |s| and |y| over segments k0.. from the sums each segment kept of its last step (no pass over the vectors)
            self._end(run32, self.seg_ss[k0:], self.seg_yy[k0:], it, crit, nsum=K - k0)
            if known is None and not self._wait(ticket):
                break
        return it

    # -- the same loop nest inside the library ------------------------------------------------------------
    def _native_plan(self):
        """The buffers of a run as a sfxm_newton_plan.  The level rows, the sums and the pinned flags belong to this driver
        and are allocated once; everything that belongs to the engine (descriptor, mesh, parameters, geometry table,
        Jacobian buffer, workspaces) is looked up again on every run - the engine may have been given new parameters,
        boundary conditions or a new mesh since the last one."""
        from . import _lib
        eng, B, dev = self.eng, self.eng.batch, self.eng.device
        P, segs = eng.problem(self.split, self.split_locs)
        K = len(segs)
        own = self._plan
        if own is None:
            own = self._plan = {"it": torch.zeros((2 * K, B), dtype=torch.int64, device=dev),
                                "crit": torch.zeros((2 * K, B), dtype=torch.float64, device=dev),
                                "run": torch.zeros((2 * K, B), dtype=torch.int32, device=dev),
                                "norms2": torch.zeros(2 * B, dtype=torch.float64, device=dev),
                                "any_host": torch.zeros(2 * K, dtype=torch.int32).pin_memory()}
        pl = _lib.NewtonPlan()
        pl.prob = ctypes.pointer(P)
        pl.x_int, pl.params, pl.param_stride = _lib.ptr(eng.x_int), _lib.ptr(eng.params), eng.param_stride
        pl.geom = _lib.ptr(eng._geom)
        pl.eps, pl.abs_tol, pl.rel_tol = float(self.eps), ABS_TOL, REL_TOL
        pl.nseg = K
        keep = [P]
        if self.fused:
            pl.solver = _lib.NEWTON_FUSED_STEP
            work = eng._scratch("solve", eng.lib.sfxm_solve_workspace_bytes(ctypes.byref(P)) // 8)
        else:
            if eng.W == 4:
                pl.solver = _lib.NEWTON_BAND4
                work = eng._scratch("band4", eng.lib.sfxm_band4_workspace_bytes(ctypes.byref(P)) // 8)
            elif B >= 2048:
                pl.solver = _lib.NEWTON_THOMAS
                work = eng._scratch("solve", eng.lib.sfxm_solve_workspace_bytes(ctypes.byref(P)) // 8)
            else:
                pl.solver = _lib.NEWTON_PCR
                work = eng._scratch("pcr", eng.lib.sfxm_pcr_workspace_bytes(ctypes.byref(P)) // 8)
            J = eng.jacobian_buffer()
            wn = eng._scratch("norms2", 2 * eng.lib.sfxm_sumsq_workspace_bytes(eng.n, B) // 8)
            pl.J, pl.work_norms, pl.r = _lib.ptr(J), _lib.ptr(wn), _lib.ptr(self.r)
            keep += [J, wn]
        keep.append(work)
        for k in range(K):
            pl.row_lo[k], pl.row_hi[k] = self.rows(k)
            if pl.solver == _lib.NEWTON_BAND4:
                pl.solve_launches[k] = 2 + max((eng.nx + 1) // 2 - 1, 1).bit_length()
            elif pl.solver == _lib.NEWTON_PCR:
                pl.solve_launches[k] = eng._pcr_launches(P, segs[k])
            else:
                pl.solve_launches[k] = 1
        pl.work_solve, pl.norms2, pl.status, pl.s = _lib.ptr(work), _lib.ptr(own["norms2"]), _lib.ptr(self.status), _lib.ptr(self.s)
        pl.it, pl.crit, pl.run = _lib.ptr(own["it"]), _lib.ptr(own["crit"]), _lib.ptr(own["run"])
        pl.seg_ss, pl.seg_yy = _lib.ptr(self.seg_ss), _lib.ptr(self.seg_yy)
        pl.any_dev, pl.any_host = _lib.ptr(self._any), own["any_host"].data_ptr()
        return pl, keep

    def _engine_key(self):
        """Identity of everything in the plan that belongs to the engine (see _native_plan)."""
        eng = self.eng
        P, _ = eng.problem(self.split, self.split_locs)
        w = eng._work
        return (id(P), eng.x_int.data_ptr(), eng.params.data_ptr(), eng.param_stride, 0 if eng._geom is None else eng._geom.data_ptr(),
                id(w), tuple(0 if w.get(k) is None else w[k].data_ptr() for k in ("J", "solve", "pcr", "band4", "norms2")))

    def _run_native(self, y, maxiter):
        from . import _lib
        key = self._engine_key() if self._plan_cache is not None else None
        if self._plan_cache is None or self._plan_cache[0] != key:
            pl, keep = self._native_plan()
            self._plan_cache = (self._engine_key(), pl, keep)
        _, pl, keep = self._plan_cache
        pl.y = _lib.ptr(y)
        level, launches = ctypes.c_int32(0), ctypes.c_int64(0)
        _lib.check(self._lib.sfxm_newton_run(ctypes.byref(pl), float(min(maxiter, 1e300)), ctypes.byref(level),
                                             ctypes.byref(launches), _lib.stream_ptr()))
        _lib.count(int(launches.value))
        return self._plan["it"][level.value].clone()

    def run(self, y, maxiter=math.inf):
        """In place on y [n, batch]; returns per-member iteration counts (int64 [batch])."""
        self.check_initial(y)
        if self.native:
            return self._run_native(y, maxiter)
        if len(self.segs) == 1:
            return self.newton_segment(y, 0, maxiter, None).clone()
        return self.split_newton(y, 0, maxiter, None).clone()
