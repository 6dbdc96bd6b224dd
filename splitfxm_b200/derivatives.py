"""Kernel-level finite-difference entry points with the reference's signatures (splitfxm/derivatives.pyx:196-281):
``Dx`` / ``D2x`` take a callable evaluated on the stencil's cells, ``dx`` / ``d2x`` take precomputed scalar values.
The callable runs on the host (as in the reference's wrappers); the stencil arithmetic runs in
csrc/sfxm_stencil_api.cu.
"""
import numpy as np
import torch

from . import _lib
from .error import SFXM
from .flux import _device
from .schemes import FDSchemes, _scheme_ids, stencil_sizes


def _derivative(values, cell_sub, scheme, order):
    if not isinstance(scheme, FDSchemes):
        raise SFXM("Unsupported scheme")
    if len(cell_sub) != stencil_sizes[scheme]:
        raise SFXM(f"Improper stencil size. Expected {stencil_sizes[scheme]}, got {len(cell_sub)}")
    lib, dev = _lib.load(), _device()
    Fv = torch.from_numpy(np.ascontiguousarray(values, dtype=np.double)).to(dev)
    xs = torch.from_numpy(np.array([cell.x() for cell in cell_sub], dtype=np.double)).to(dev)
    nv = Fv.shape[1]
    out = torch.empty(nv, dtype=torch.float64, device=dev)
    _lib.check(lib.sfxm_fd_derivative(_scheme_ids[scheme], order, nv, Fv.data_ptr(), xs.data_ptr(), out.data_ptr(),
                                      _lib.stream_ptr()))
    return out.cpu().numpy()


def _call(fn, cell_sub):
    return np.array([fn(cell.values()) for cell in cell_sub], dtype=np.double)


def Dx(F, cell_sub, scheme):
    """First derivative of F(values) over the stencil (derivatives.pyx:196-216)."""
    return _derivative(_call(F, cell_sub), cell_sub, scheme, 1)


def D2x(D, cell_sub, scheme):
    """Second derivative of D(values) over the stencil (derivatives.pyx:240-260)."""
    return _derivative(_call(D, cell_sub), cell_sub, scheme, 2)


def dx(values, cell_sub, scheme):
    """First derivative of precomputed scalar values (derivatives.pyx:218-238)."""
    return float(_derivative(np.asarray(values, dtype=np.double).reshape(-1, 1), cell_sub, scheme, 1)[0])


def d2x(values, cell_sub, scheme):
    """Second derivative of precomputed scalar values (derivatives.pyx:262-281)."""
    return float(_derivative(np.asarray(values, dtype=np.double).reshape(-1, 1), cell_sub, scheme, 2)[0])
