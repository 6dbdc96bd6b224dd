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


class EnsembleSimulation:
    def __init__(self, d, model, params, bcs, scheme, method="FVM", device=None):
        """d: Domain (mesh, ghosts, components); model: 'ad' | 'euler_prim' | 'euler_cons'; params: dict name ->
        scalar or array[batch]; bcs: reference-style dict shared by all members."""
        if model not in _MODELS:
            raise SFXM("unknown device model " + str(model))
        if (method == "FVM") != isinstance(scheme, FVSchemes):
            raise SFXM("Scheme family does not match the numerical method")
        nbl, nbr = d.nb(btype.LEFT), d.nb(btype.RIGHT)
        if stencil_sizes.get(scheme) != nbl + nbr + 1:
            raise SFXM("Stencil size does not match boundary conditions")
        model_id, names = _MODELS[model]
        cols = [np.atleast_1d(np.asarray(params[k], dtype=np.float64)) for k in names]
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

    def steady_state(self, Y, split=False, split_locs=None, dt0=0.0, dtmax=1.0, armijo=False, bounds=None,
                     maxiter=math.inf):
        """-> (Y_final [n, batch] device tensor, iterations int64[batch])."""
        y = self._state(Y).clone()
        drv = NewtonDriver(self.eng, split, split_locs, dt0, dtmax, armijo, bounds)
        its = drv.run(y, maxiter)
        return y, its

    def evolve_euler(self, Y, t_diff, max_step, split=False, split_locs=None):
        y, _, _ = euler(self.eng, self._state(Y).clone(), t_diff, max_step, split, split_locs)
        return y
