"""Finite-volume transport equation (reference: splitfxm/equations/fv_transport.py:4-59).

Same constructor (F, D, S, dFdU callables) and the same ``residuals(cell_sub, scheme, scheme_opts)``: the callables are
evaluated on the three stencil cells on the host, the arithmetic - Lax-Friedrichs / central / upwind fluxes, central
diffusion fluxes and rhs = (S - (1/dx)(Fe - Fw)) + (1/dx)(DFe - DFw) - is one launch of sfxm_fv_cell_residual.  The
mesh-wide hot path does not come through here: csrc/sfxm_device.cuh cell_residual evaluates device functors."""
import numpy as np
import torch

from .. import _lib
from ..error import SFXM
from ..schemes import FVSchemes, _scheme_ids, stencil_sizes
from ._stencil import device, source_values, stencil_values, stencil_x


class FVTransportEquation:
    method = "FVM"

    def __init__(self, F, D, S, dFdU):
        self.F = F
        self.D = D
        self.S = S
        self.dFdU = dFdU

    def residuals(self, cell_sub, scheme, scheme_opts={}):
        if not isinstance(scheme, FVSchemes) or scheme not in stencil_sizes:
            raise SFXM("Unsupported scheme")
        if len(cell_sub) != stencil_sizes[scheme]:
            raise SFXM(f"Improper stencil size. Expected {stencil_sizes[scheme]}, got {len(cell_sub)}")
        scheme_opts.get("limiter")                 # read and unused, as in the reference (fv_transport.py:55)
        lib, dev = _lib.load(), device()
        uc = cell_sub[len(cell_sub) // 2].values()
        sigma, eig = 0.0, None
        if scheme == FVSchemes.LAX_FRIEDRICHS:
            if self.dFdU is None:
                raise SFXM("dFdU is required for determining spectral radius")
            sigma = float(np.max(np.abs(np.linalg.eigvals(self.dFdU(uc)))))          # flux.pyx:98-99
        elif scheme == FVSchemes.UPWIND:
            if self.dFdU is None:
                raise SFXM("dFdU is required for determining upwind direction.")
            A = np.asarray(self.dFdU(uc), dtype=np.double)
            if np.count_nonzero(A - np.diag(np.diagonal(A))):
                raise SFXM("UPWIND is supported for a diagonal dFdU (the general case depends on LAPACK's eigenvector order)")
            eig = torch.from_numpy(np.ascontiguousarray(np.diagonal(A))).to(dev)
        Fv = stencil_values(self.F, cell_sub, dev)
        Dv = stencil_values(self.D, cell_sub, dev)
        uv = stencil_values(lambda v: v, cell_sub, dev)
        nv = Fv.shape[1]
        out = torch.empty(nv, dtype=torch.float64, device=dev)
        xs, Sv = stencil_x(cell_sub, dev), source_values(self.S, cell_sub, nv, dev)     # named: alive until the launch
        _lib.check(lib.sfxm_fv_cell_residual(_scheme_ids[scheme], nv, Fv.data_ptr(), Dv.data_ptr(), uv.data_ptr(),
                                             xs.data_ptr(), Sv.data_ptr(), sigma, _lib.ptr(eig), out.data_ptr(),
                                             _lib.stream_ptr()))
        _lib.count()
        return out.cpu().numpy()
