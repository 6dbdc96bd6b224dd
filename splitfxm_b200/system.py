"""System: model + scheme (+ options) (reference: splitfxm/system.py:6-70).

``residuals(d)`` of the reference walks the interior in Python, one Cython call per cell; here it is ONE launch of the
residual kernel (csrc/sfxm_kernels.cuh) over the Domain's device arrays, with the ghost cells taken as the Domain
holds them - the reference's method does not apply boundary conditions either (its caller does, simulation.py:174-177).
A model without a device functor falls back to the per-cell equation objects, whose stencil arithmetic is a kernel too
(equations/*.py -> sfxm_fv_cell_residual / sfxm_fd_cell_residual)."""
import numpy as np

from .bc import describe_bcs
from .constants import btype
from .error import SFXM
from .schemes import FVSchemes, _scheme_ids


class System:
    def __init__(self, model, scheme, scheme_opts):
        self._model = model
        self._scheme = scheme
        self._scheme_opts = scheme_opts

    def residuals(self, d):
        """-> ndarray [nx, nv]: the right-hand side of every interior cell (system.py:28-70)."""
        dev = getattr(self._model, "_device", None)
        if dev is None:
            # opaque Python callables: the reference's loop, each cell's arithmetic on the device
            cells = d.cells()
            nbl, nbr = d.nb(btype.LEFT), d.nb(btype.RIGHT)
            eq = self._model.equation()
            return np.vstack([eq.residuals([cells[i + o] for o in range(-nbl, nbr + 1)], self._scheme, self._scheme_opts)
                              for i in range(d.ilo(), d.ihi() + 1)])
        import torch
        from .engine import Engine
        if not torch.cuda.is_available():
            raise SFXM("splitfxm_b200 needs a CUDA device (there is no CPU fallback)")
        if (dev.method == "FVM") != isinstance(self._scheme, FVSchemes) or self._scheme not in _scheme_ids:
            raise SFXM("Unsupported scheme")
        nbl, nbr, nx = d.nb(btype.LEFT), d.nb(btype.RIGHT), d._nx
        dx, dU = d._device()
        eng = Engine(dev.model_id, _scheme_ids[self._scheme], d.num_components(), nbl, nbr, dx[nbl:nbl + nx], d.xmin(),
                     d.xmax(), dev.params, batch=1)
        eng.set_bcs(lambda P: describe_bcs(P, {}, d))          # no BC: the ghost cells are what the Domain holds
        y = eng.block_to_flat(dU[nbl:nbl + nx])
        return eng.residual(y).reshape(nx, eng.nv).cpu().numpy()
