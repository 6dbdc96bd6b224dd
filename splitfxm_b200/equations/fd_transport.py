"""Finite-difference transport equation (reference: splitfxm/equations/fd_transport.py:4-51).

Same constructor (F, D, S callables) and ``residuals(cell_sub, scheme, scheme_opts)``: rhs = (S - Dx(F)) + D2x(D) with
the CENTRAL (3 cells) or RIGHT_BIAS (4 cells) stencils of derivatives.pyx:26-125, one launch of sfxm_fd_cell_residual on
the values of the callables."""
import torch

from .. import _lib
from ..error import SFXM
from ..schemes import FDSchemes, _scheme_ids, stencil_sizes
from ._stencil import device, source_values, stencil_values, stencil_x


class FDTransportEquation:
    method = "FDM"

    def __init__(self, F, D, S):
        self.F = F
        self.D = D
        self.S = S

    def residuals(self, cell_sub, scheme, scheme_opts={}):
        if not isinstance(scheme, FDSchemes):
            raise SFXM("Unsupported scheme")
        if len(cell_sub) != stencil_sizes[scheme]:
            raise SFXM(f"Improper stencil size. Expected {stencil_sizes[scheme]}, got {len(cell_sub)}")
        lib, dev = _lib.load(), device()
        Fv = stencil_values(self.F, cell_sub, dev)
        Dv = stencil_values(self.D, cell_sub, dev)
        nv = Fv.shape[1]
        out = torch.empty(nv, dtype=torch.float64, device=dev)
        xs, Sv = stencil_x(cell_sub, dev), source_values(self.S, cell_sub, nv, dev)     # named: alive until the launch
        _lib.check(lib.sfxm_fd_cell_residual(_scheme_ids[scheme], nv, Fv.data_ptr(), Dv.data_ptr(), xs.data_ptr(),
                                             Sv.data_ptr(), out.data_ptr(), _lib.stream_ptr()))
        _lib.count()
        return out.cpu().numpy()
