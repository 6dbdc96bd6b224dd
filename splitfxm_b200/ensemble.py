"""EnsembleSimulation: many independent members (parameter sweeps, many initial states) on one mesh.

The reference has no batching API (one ``Simulation`` is one Python object graph, simulation.py:52-75); this is
the B200-native way to run BASELINE.json's 65,536-member sweep: every member is one lane of the member-fastest
layout, all members advance in the same kernel launches, and members shard across GPUs with no communication.
State lives in device tensors ``[n, batch]`` (flat index major, member minor).
"""
import math

import numpy as np
import torch

from . import _lib
from .bc import describe_bcs
from .constants import btype
from .engine import Engine
from .error import SFXM
from .newton import NewtonDriver
from .ode import euler
from .schemes import FVSchemes, _scheme_ids, stencil_sizes

_MODELS = {"ad": (_lib.MODEL_AD, ("c", "nu")), "euler_prim": (_lib.MODEL_EULER_PRIM, ("gamma",)),
           "euler_cons": (_lib.MODEL_EULER_CONS, ("gamma",))}


def resolve_model(model, params):
    """-> (model_id, [one scalar or array[batch] per device parameter]).  model: a built-in name ('ad' | 'euler_prim' |
    'euler_cons', params: dict name -> value) or a DeviceModel, e.g. from custom.compile_model (params: None for the
    DeviceModel's own values, or a sequence / dict with one entry per parameter of the functor, in its order)."""
    if hasattr(model, "model_id"):
        if params is None:
            return model.model_id, list(model.params)
        cols = list(params.values()) if isinstance(params, dict) else list(params)
        if len(cols) != len(model.params):
            raise SFXM("one value (or array over the members) per parameter of the device model is needed")
        return model.model_id, cols
    if model not in _MODELS:
        raise SFXM("unknown device model " + str(model))
    model_id, names = _MODELS[model]
    return model_id, [params[k] for k in names]


class EnsembleSimulation:
    def __init__(self, d, model, params, bcs, scheme, method="FVM", device=None):
        """d: Domain (mesh, ghosts, components); model: 'ad' | 'euler_prim' | 'euler_cons' or a DeviceModel (custom
        models, custom.compile_model); params: dict name -> scalar or array[batch] (see resolve_model); bcs:
        reference-style dict shared by all members."""
        model_id, cols = resolve_model(model, params)
        if (method == "FVM") != isinstance(scheme, FVSchemes):
            raise SFXM("Scheme family does not match the numerical method")
        nbl, nbr = d.nb(btype.LEFT), d.nb(btype.RIGHT)
        if stencil_sizes.get(scheme) != nbl + nbr + 1:
            raise SFXM("Stencil size does not match boundary conditions")
        cols = [np.atleast_1d(np.asarray(c, dtype=np.float64)) for c in cols]
        batch = max(c.shape[0] for c in cols)
        table = np.stack([np.broadcast_to(c, (batch,)) for c in cols])
        self._d, self._bcs = d, bcs
        self.batch = batch
        self.eng = Engine(model_id, _scheme_ids[scheme], d.num_components(), nbl, nbr, d.interior_x(), d.xmin(),
                          d.xmax(), table, batch=batch, device=device)
        self.eng.set_bcs(lambda P: describe_bcs(P, bcs, d))

    def _state(self, Y):
        if isinstance(Y, torch.Tensor):
            if Y.shape != (self.eng.n, self.batch) or Y.dtype != torch.float64:
                raise SFXM("state must be float64 [n, batch]")
            return Y.to(self.eng.device).contiguous()
        return self.eng.to_device(Y)

    def residual(self, Y, split=False, split_locs=None, out=None):
        return self.eng.residual(self._state(Y), out, split, split_locs)

    def jacobian(self, Y, split=False, split_locs=None, epsilon=1e-8, out=None):
        """Block-banded Jacobians [nx, W, nv, nv, batch] (include/sfxm_b200.h)."""
        return self.eng.jacobian(self._state(Y), split, split_locs, epsilon, J=out)

    def jacobian_host(self, Y_host, out_host, split=False, split_locs=None, epsilon=1e-8, slot=0):
        """Host buffers in, host buffers out, asynchronously: Y_host [n, batch] (pinned) -> out_host
        [nx, W, L, batch] (pinned) with the L same-segment entries of every block (`eng.live_entries`).

        The upload, the assembly and the packing run on one stream, the download on a second one; with two slots
        (`slot` 0 / 1, each with its own device buffers) consecutive calls overlap the download of one step with the
        upload and assembly of the next.  Call `sync_host()` before reading `out_host`."""
        eng = self.eng
        if not hasattr(self, "_slots"):
            self._slots = {}
            self._copy_stream = torch.cuda.Stream(device=eng.device)
        st = self._slots.get(slot)
        if st is None:
            L = len(eng.live_entries(split, split_locs))
            st = {"Y": torch.empty((eng.n, eng.batch), dtype=torch.float64, device=eng.device),
                  "J": torch.zeros((eng.nx, eng.W, eng.nv, eng.nv, eng.batch), dtype=torch.float64, device=eng.device),
                  "Jc": torch.empty((eng.nx, eng.W, L, eng.batch), dtype=torch.float64, device=eng.device),
                  "done": torch.cuda.Event(), "free": torch.cuda.Event()}
            st["free"].record()
            self._slots[slot] = st
        main = torch.cuda.current_stream()
        main.wait_event(st["free"])                     # the previous download from this slot has finished
        st["Y"].copy_(Y_host, non_blocking=True)
        eng.jacobian(st["Y"], split, split_locs, epsilon, J=st["J"])
        eng.pack_jacobian(st["J"], split, split_locs, out=st["Jc"])
        st["done"].record(main)
        with torch.cuda.stream(self._copy_stream):
            self._copy_stream.wait_event(st["done"])
            out_host.copy_(st["Jc"], non_blocking=True)
            st["free"].record(self._copy_stream)
        return out_host

    def sync_host(self):
        if hasattr(self, "_ss_up"):
            self._ss_up.synchronize()
        if hasattr(self, "_copy_stream"):
            self._copy_stream.synchronize()
        torch.cuda.current_stream().synchronize()

    def steady_state(self, Y, split=False, split_locs=None, dt0=0.0, dtmax=1.0, armijo=False, bounds=None,
                     maxiter=math.inf):
        """-> (Y_final [n, batch] device tensor, iterations int64[batch])."""
        y = self._state(Y).clone()
        # the driver (work vectors, per-level counters, events) is kept between calls with the same options
        key = (bool(split), tuple(split_locs) if split_locs else None, float(dt0), float(dtmax), bool(armijo))
        drv = getattr(self, "_driver", None)
        if bounds is not None or drv is None or self._driver_key != key:
            drv = NewtonDriver(self.eng, split, split_locs, dt0, dtmax, armijo, bounds)
            if bounds is None:
                self._driver, self._driver_key = drv, key
        its = drv.run(y, maxiter)
        return y, its

    def steady_state_host(self, Y_host, out_host, its_host, split=False, split_locs=None, dt0=0.0, dtmax=1.0,
                          armijo=False, bounds=None, maxiter=math.inf, slot=None, next_Y_host=None):
        """Host buffers in, host buffers out: Y_host [n, batch] (pinned) is uploaded, every member is Newton-solved on
        the device, and the final states / iteration counts are downloaded into out_host [n, batch] / its_host [batch]
        (pinned).

        slot = None: synchronous - returns when the results are in host memory.
        slot = 0 / 1: pipelined over three streams (upload, solve, download) and two device state buffers: the download
        of one step overlaps the upload and the solve of the next, and `next_Y_host` (the input of the NEXT call, which
        must use the other slot) is uploaded while this call's Newton iterations run.  Call `sync_host()` before reading
        `out_host` / `its_host`."""
        eng = self.eng
        key = (bool(split), tuple(split_locs) if split_locs else None, float(dt0), float(dtmax), bool(armijo), bounds is None)
        drv = getattr(self, "_ss_driver", None)
        if drv is None or self._ss_key != key or bounds is not None:
            drv = NewtonDriver(eng, split, split_locs, dt0, dtmax, armijo, bounds)
            self._ss_driver, self._ss_key = drv, key
        if slot is None:
            if not hasattr(self, "_ss_state"):
                self._ss_state = torch.empty((eng.n, eng.batch), dtype=torch.float64, device=eng.device)
            y = self._ss_state
            y.copy_(Y_host, non_blocking=True)
            its = drv.run(y, maxiter)
            out_host.copy_(y, non_blocking=True)
            its_host.copy_(its, non_blocking=True)
            torch.cuda.current_stream().synchronize()
            return out_host, its_host
        if not hasattr(self, "_ss_slots"):
            self._ss_slots = {}
            self._ss_up = torch.cuda.Stream(device=eng.device)
            if not hasattr(self, "_copy_stream"):
                self._copy_stream = torch.cuda.Stream(device=eng.device)

        def slot_state(k):
            st = self._ss_slots.get(k)
            if st is None:
                st = {"y": torch.empty((eng.n, eng.batch), dtype=torch.float64, device=eng.device),
                      "its": torch.empty(eng.batch, dtype=torch.int64, device=eng.device),
                      "up": torch.cuda.Event(), "done": torch.cuda.Event(), "free": torch.cuda.Event(), "loaded": None}
                st["free"].record()
                self._ss_slots[k] = st
            return st

        def upload(st, src):
            with torch.cuda.stream(self._ss_up):
                self._ss_up.wait_event(st["free"])       # the previous download from this slot has finished
                st["y"].copy_(src, non_blocking=True)
                st["up"].record(self._ss_up)
            st["loaded"] = src

        st = slot_state(slot)
        if st["loaded"] is not Y_host:
            upload(st, Y_host)
        if next_Y_host is not None:
            upload(slot_state(1 - slot), next_Y_host)    # overlaps this call's Newton iterations
        main = torch.cuda.current_stream()
        main.wait_event(st["up"])
        its = drv.run(st["y"], maxiter)
        st["its"].copy_(its)
        st["done"].record(main)
        st["loaded"] = None
        with torch.cuda.stream(self._copy_stream):
            self._copy_stream.wait_event(st["done"])
            out_host.copy_(st["y"], non_blocking=True)
            its_host.copy_(st["its"], non_blocking=True)
            st["free"].record(self._copy_stream)
        return out_host, its_host

    def evolve_euler(self, Y, t_diff, max_step, split=False, split_locs=None):
        y, _, _ = euler(self.eng, self._state(Y).clone(), t_diff, max_step, split, split_locs)
        return y
