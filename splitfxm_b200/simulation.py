"""Simulation: the reference's solver front end (splitfxm/simulation.py:30-526) over the CUDA engine.

Same constructor, methods, argument meaning and error behaviour as the reference.  Each public call uploads
the Domain's interior once, runs entirely on the device (ghost fill, residual, Jacobian, linear solves, time
stepping - csrc/*.cu) and writes the final state, ghosts included, back into the Domain.
"""
import math

import numpy as np
import torch
from scipy.sparse import lil_matrix

from . import _lib
from .bc import apply_BC, describe_bcs
from .constants import btype
from .domain import Domain
from .engine import Engine
from .error import SFXM
from .initialize import set_initial_condition
from .model import Model
from .newton import NewtonDriver
from .ode import RK45Stepper, euler
from .refine import Refiner
from .schemes import FDSchemes, FVSchemes, _scheme_ids, stencil_sizes
from .simhelper import DeviceBounds, array_list_reshape, create_bound_events, report_bounds
from .system import System

JAC_REQUIRED = ["Radau", "BDF"]


def _lil_from_coo(rows, cols, vals, n):
    """lil_matrix with the given (unique) entries, filled row list by row list - several times cheaper than
    coo_matrix(...).tolil() for the few thousand entries of a banded Jacobian."""
    order = np.lexsort((cols, rows))
    rows, cols, vals = rows[order], cols[order], vals[order]
    counts = np.bincount(rows, minlength=n)
    cl, vl = cols.tolist(), vals.tolist()
    L = lil_matrix((n, n))
    at = 0
    lrows, ldata = L.rows, L.data
    for i, c in enumerate(counts.tolist()):
        if c:
            lrows[i] = cl[at:at + c]
            ldata[i] = vl[at:at + c]
            at += c
    return L


class Simulation:
    def __init__(self, d: Domain, m: Model, ics: dict, bcs: dict, scheme, scheme_opts: dict = {}, ss: dict = {}):
        self._d = d
        self._s = System(m, scheme, scheme_opts)
        self._r = Refiner()
        self._bcs = bcs
        self._ss = ss          # stored and never read by the reference (simulation.py:62); "maxiter" is honoured here

        for c, ictype in ics.items():
            set_initial_condition(self._d, c, ictype)
        for c, bctype in self._bcs.items():
            apply_BC(self._d, c, bctype)

        boundary_cells = self._d.nb(btype.LEFT) + self._d.nb(btype.RIGHT)
        if stencil_sizes.get(self._s._scheme) != boundary_cells + 1:
            raise SFXM("Stencil size does not match boundary conditions")
        self._eng = None
        self._eng_key = self._eng_rev = self._bc_sig = None
        self._rk = None
        self._nd = None

    # -- device engine -----------------------------------------------------------------------------------
    def _engine(self):
        """Engine for the Domain's current mesh.  The Domain's device copy is the engine's input: nothing is uploaded
        unless host code touched the Domain since the last call (domain.py: _device)."""
        d, m = self._d, self._s._model
        dev = getattr(m, "_device", None)
        if dev is None:
            raise SFXM("Model has no device functor: the hot path runs on the GPU only (no CPU fallback). "
                       "Use splitfxm_b200.models.* or attach a DeviceModel to the model.")
        scheme = self._s._scheme
        if (dev.method == "FVM") != isinstance(scheme, FVSchemes):
            raise SFXM("Scheme family does not match the model's numerical method")
        if scheme not in _scheme_ids:
            raise SFXM("Unsupported scheme")
        if not torch.cuda.is_available():
            raise SFXM("splitfxm_b200 needs a CUDA device (there is no CPU fallback)")
        nbl, nbr, nx = d.nb(btype.LEFT), d.nb(btype.RIGHT), d._nx
        dx, _ = d._device()
        key = (nx, d.num_components(), nbl, nbr, _scheme_ids[scheme], dev.model_id)
        if self._eng is None or self._eng_key != key:
            self._eng = Engine(dev.model_id, _scheme_ids[scheme], d.num_components(), nbl, nbr, dx[nbl:nbl + nx],
                               d.xmin(), d.xmax(), dev.params, batch=1)
            self._eng_key, self._eng_rev, self._bc_sig = key, d._mesh_rev, None
        else:
            if self._eng_rev != d._mesh_rev:
                self._eng.set_mesh(dx[nbl:nbl + nx])
                self._eng_rev, self._bc_sig = d._mesh_rev, None
            self._eng.set_params(dev.params)
        # the BC descriptor is rebuilt when the bcs dict changed, and always when some variable has no BC (its ghost
        # cells are whatever the Domain holds at the time of the call)
        sig = repr(self._bcs) if len(self._bcs) == d.num_components() else None
        if sig is None or sig != self._bc_sig:
            self._eng.set_bcs(lambda P: describe_bcs(P, self._bcs, d))
            self._bc_sig = sig
        return self._eng

    def _state(self, eng, split, split_locs):
        """The Domain's interior as the flat device vector [n, 1] (listify_interior, domain.py:276-312) - no host copy."""
        d = self._d
        nbl = d.nb(btype.LEFT)
        _, dU = d._device()
        return eng.block_to_flat(dU[nbl:nbl + d._nx], split, split_locs)

    def _store(self, eng, y, split, split_locs, ghost_state=None):
        """Device state -> Domain: initialize_from_list + apply_BC (simulation.py:171-175), on the device.

        ``ghost_state``: the state of the last right-hand-side evaluation.  The reference's evolve() ends with
        initialize_from_list (simulation.py:472), which rewrites the interior only, so the ghost cells it leaves
        behind are those of the last RHS call - reproduced here."""
        d = self._d
        x_all, U_all = eng.fill_domain(y if ghost_state is None else ghost_state, split, split_locs)
        U_all = U_all.reshape(x_all.shape[0], eng.nv)
        if ghost_state is not None:
            nbl = d.nb(btype.LEFT)
            U_all[nbl:nbl + eng.nx] = eng.flat_to_block(y, split, split_locs)
        d._adopt_device(x_all, U_all)      # ghost positions may move (bc.py:56-151); the interior mesh does not

    # -- list related methods (simulation.py:81-264) -----------------------------------------------------
    def get_shape_from_list(self, l):
        nv = self._d.num_components()
        if len(l) % nv != 0:
            raise SFXM("List length not aligned with interior size")
        return len(l) // nv, nv

    @staticmethod
    def _check_listify(nv, split, split_locs):
        """The errors Domain.listify_interior raises (domain.py:276-312) - evolve / steady_state start with it."""
        if split:
            if not split_locs:
                raise SFXM("Split locations must be specified when splitting is enabled")
            if sorted(set(split_locs))[-1] > nv:
                raise SFXM("Split locations exceed the number of columns in the interior values")

    def _check_split(self, nv, split, split_locs, msg_none, msg_range):
        if split:
            if split_locs is None:
                raise SFXM(msg_none)
            locs = sorted(set(split_locs))
            if locs[-1] > nv or locs[0] < 0:
                raise SFXM(msg_range)

    def initialize_from_list(self, l, split=False, split_locs=None):
        num_points, nv = self.get_shape_from_list(l)
        if split:
            self._check_split(nv, split, split_locs, "Split location must be specified in this case",
                              "Split locations must be between 0 and nv-1")
            split_locs = sorted(set(split_locs))
            l_split = np.split(np.asarray(l), [i * num_points for i in split_locs])
            block = np.hstack([array_list_reshape(segment, (num_points, -1)) for segment in l_split])
        else:
            block = array_list_reshape(l, (num_points, nv))
        if num_points != self._d._nx:
            raise SFXM("List length not aligned with interior size")
        self._d.interior_U()[:, :] = block

    def get_residuals_from_list(self, l, split=False, split_locs=None):
        num_points, nv = self.get_shape_from_list(l)
        self._check_split(nv, split, split_locs, "Split location must be specified in this case",
                          "Split locations must be between 0 and nv-1")
        eng = self._engine()
        y = eng.to_device(np.asarray(l, dtype=np.float64))
        r = eng.residual(y, None, split, split_locs)
        self._store(eng, y, split, split_locs)
        return r.cpu().numpy().reshape(-1)

    def extend_bounds(self, bounds, num_points, nv, split=False, split_locs=None):
        if bounds is None:
            return None
        if len(bounds) != 2:
            raise SFXM("Bounds must be a list of 2 lists")
        if len(bounds[0]) != nv or len(bounds[1]) != nv:
            raise SFXM("Each list in bounds must be of length - number of variables")
        if not split:
            return [bounds[0] * num_points, bounds[1] * num_points]
        if split_locs is None:
            raise SFXM("split_locs must be provided if split is True")
        split_locs = sorted(set(split_locs))
        if split_locs[-1] > nv or split_locs[0] < 0:
            raise SFXM("Split locations must be between 0 and nv-1")
        lower = np.concatenate([np.tile(seg, num_points) for seg in np.split(bounds[0], split_locs)]).tolist()
        upper = np.concatenate([np.tile(seg, num_points) for seg in np.split(bounds[1], split_locs)]).tolist()
        return [lower, upper]

    # -- solution related methods ------------------------------------------------------------------------
    def jacobian(self, l, split=False, split_locs=None, epsilon=1e-8):
        """Finite-difference Jacobian as scipy.sparse.lil_matrix (simulation.py:270-394)."""
        num_points, nv = self.get_shape_from_list(l)
        if split and split_locs is None:
            raise SFXM("split_locs must be provided if split is True")
        self._check_split(nv, split, split_locs, "Split location must be specified in this case",
                          "Split locations must be between 0 and nv-1")
        eng = self._engine()
        y = eng.to_device(np.asarray(l, dtype=np.float64))
        J = eng.jacobian(y, split, split_locs, epsilon)
        self._store(eng, y, split, split_locs)
        Jh = J.cpu().numpy().reshape(eng.nx, eng.W, nv, nv)
        rows, cols, vals = eng.jacobian_coo(Jh, split, split_locs)
        return _lil_from_coo(rows, cols, vals, num_points * nv)

    def evolve(self, t_diff: float, split=False, split_locs=None, method='RK45', rtol=1e-3, atol=1e-6,
               max_step=np.inf, bounds=None):
        """Advance the Domain by t_diff (simulation.py:396-472).  'Euler' and 'RK45' run on the device; like the
        reference, rtol/atol are accepted but scipy's defaults are what is used (they are never forwarded,
        simulation.py:468-469)."""
        num_points, nv = self._d._nx, self._d.num_components()
        self._check_listify(nv, split, split_locs)
        ext_bounds = self.extend_bounds(bounds, num_points, nv, split, split_locs)
        eng = self._engine()
        y = self._state(eng, split, split_locs)
        ghost_state = None
        if method == "Euler":
            on_step = None
            if bounds is not None:
                checker = DeviceBounds(ext_bounds, eng.device)       # compared on the device, offenders only come back

                def on_step(yd, t):
                    checker.report(yd, t)
            y, ghost_state, _ = euler(eng, y, t_diff, max_step, split, split_locs, on_step)
        elif method == "RK45":
            on_step = None
            if bounds is not None:
                # the reference hands one event per unknown to solve_ivp (simhelper.py:31-88); scipy evaluates them at the
                # start and after every accepted step, and each evaluation prints the warnings - reproduced per step (the
                # extra evaluations of scipy's event root-refinement are not)
                checker = DeviceBounds(ext_bounds, eng.device)

                def on_step(yd, t):
                    checker.report(yd, t)
            # the stepper (stage vectors, control block, captured CUDA graph of one attempted step) is kept while the
            # engine, its boundary conditions and parameters stay what they were when the graph was captured
            dev_m = self._s._model._device
            key = (id(eng), self._eng_rev, self._bc_sig, tuple(dev_m.params), bool(split), tuple(split_locs) if split_locs else None,
                   float(max_step)) if (bounds is None and self._bc_sig is not None) else None
            stepper = self._rk[1] if (key is not None and self._rk is not None and self._rk[0] == key) else None
            if stepper is None:
                stepper = RK45Stepper(eng, split, split_locs, 1e-3, 1e-6, max_step, on_step=on_step)
                self._rk = (key, stepper) if key is not None else None
            stepper.nfev = 0
            y = stepper.integrate(y, t_diff)
            ghost_state = stepper.last_rhs_state
        else:
            # Any other scipy method (RK23, DOP853, LSODA, Radau, BDF ...): scipy.integrate.solve_ivp drives, exactly as
            # in the reference (simulation.py:466-469), with the right-hand side and - for the implicit methods - the
            # constant Jacobian at y0 (simulation.py:441-443) evaluated on the device.  scipy's own step control and
            # linear algebra stay scipy's, as they are in the reference.
            from scipy.integrate import solve_ivp
            last = {}
            y0 = y.cpu().numpy().reshape(-1)

            def f(_, yv):
                yd = eng.to_device(np.asarray(yv, dtype=np.float64))
                last["y"] = yd
                return eng.residual(yd, None, split, split_locs).cpu().numpy().reshape(-1)

            jac = None
            solver = method
            if method == "BDF" and eng.W == 3:
                # scipy's step control as in the reference; the Newton systems (I - c J) d = b on the device (ode.device_bdf)
                from scipy.sparse import identity
                from .ode import device_bdf
                solver = device_bdf(eng, y, split, split_locs)
                jac = identity(y0.shape[0], format="csc")      # placeholder of the right shape: replaced by the device J
            elif method == "Radau" and eng.W == 3:
                # the real and the complex shifted systems of every simplified-Newton iteration on the device
                # (ode.device_radau); segments wider than two variables keep scipy's host factorisation
                from scipy.sparse import identity
                from .ode import device_radau
                solver = device_radau(eng, y, split, split_locs)
                if solver is None:
                    solver = method
                    jac = self.jacobian(y0, split, split_locs)
                    eng = self._engine()
                else:
                    jac = identity(y0.shape[0], format="csc")
            elif method in JAC_REQUIRED:
                jac = self.jacobian(y0, split, split_locs)
                eng = self._engine()
            events = create_bound_events(y0, ext_bounds)
            sol = solve_ivp(f, (0, t_diff), y0, method=solver, t_eval=[t_diff], jac=jac, max_step=max_step, events=events)
            y = eng.to_device(np.ascontiguousarray(sol.y.reshape(-1)))
            ghost_state = last.get("y")
        self._store(eng, y, split, split_locs, ghost_state)

    def steady_state(self, split=False, split_locs=None, sparse=True, dt0=0.0, dtmax=1.0, armijo=False, bounds=None):
        """Newton / Split-Newton to steady state (simulation.py:474-526); returns the iteration count."""
        num_points, nv = self._d._nx, self._d.num_components()
        self._check_listify(nv, split, split_locs)
        ext_bounds = self.extend_bounds(bounds, num_points, nv, split, split_locs)
        if split:
            self._check_split(nv, split, split_locs, "Split location must be specified in this case",
                              "Split locations must be between 0 and nv-1")
        eng = self._engine()
        y = self._state(eng, split, split_locs)
        # the driver (work vectors, per-level counters, events) is kept while the engine and the options stay the same
        key = (id(eng), self._eng_rev, bool(split), tuple(split_locs) if split_locs else None, float(dt0), float(dtmax),
               bool(armijo)) if bounds is None else None
        drv = self._nd[1] if (key is not None and self._nd is not None and self._nd[0] == key) else None
        if drv is None:
            drv = NewtonDriver(eng, split, split_locs, dt0, dtmax, armijo, ext_bounds)
            self._nd = (key, drv) if key is not None else None
        maxiter = self._ss.get("maxiter", math.inf) if isinstance(self._ss, dict) else math.inf
        it = drv.run(y, maxiter)
        self._store(eng, y, split, split_locs)
        return int(it[0])
