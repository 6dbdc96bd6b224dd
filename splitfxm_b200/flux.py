"""Kernel-level finite-volume entry points with the reference's signatures (splitfxm/flux.pyx:42-124, 302-327).

``F``, ``D`` and ``dFdU`` are Python callables, so - exactly as in the reference's wrappers - they are evaluated on
the stencil's cells on the host; the stencil arithmetic itself runs in csrc/sfxm_stencil_api.cu.
"""
import numpy as np
import torch

from . import _lib
from .error import SFXM
from .schemes import FVSchemes, _scheme_ids, stencil_sizes


def _device():
    if not torch.cuda.is_available():
        raise SFXM("splitfxm_b200 needs a CUDA device (there is no CPU fallback)")
    return torch.device("cuda", torch.cuda.current_device())


def _check_stencil(cell_sub, scheme):
    if not isinstance(scheme, FVSchemes) or scheme not in stencil_sizes:
        raise SFXM("Unsupported scheme")
    if len(cell_sub) != stencil_sizes[scheme]:
        raise SFXM(f"Improper stencil size. Expected {stencil_sizes[scheme]}, got {len(cell_sub)}")


def _values(fn, cell_sub):
    return np.ascontiguousarray(np.array([fn(cell.values()) for cell in cell_sub], dtype=np.double))


def fluxes(F, cell_sub, scheme, dFdU=None, limiter=None):
    """West and east fluxes of the centre cell of a 3-cell stencil (flux.pyx:42-124).  `limiter` is accepted and
    unused, as in the reference."""
    _check_stencil(cell_sub, scheme)
    lib, dev = _lib.load(), _device()
    sigma, eig = 0.0, None
    uc = cell_sub[1].values()
    if scheme == FVSchemes.LAX_FRIEDRICHS:
        if dFdU is None:
            raise SFXM("dFdU is required for determining spectral radius")
        sigma = float(np.max(np.abs(np.linalg.eigvals(dFdU(uc)))))            # flux.pyx:98-99
    elif scheme == FVSchemes.UPWIND:
        if dFdU is None:
            raise SFXM("dFdU is required for determining upwind direction.")
        A = np.asarray(dFdU(uc), dtype=np.double)
        if np.count_nonzero(A - np.diag(np.diagonal(A))):
            raise SFXM("UPWIND is supported for a diagonal dFdU (the general case depends on LAPACK's eigenvector order)")
        eig = torch.from_numpy(np.ascontiguousarray(np.diagonal(A))).to(dev)
    Fv = torch.from_numpy(_values(F, cell_sub)).to(dev)
    uv = torch.from_numpy(_values(lambda v: v, cell_sub)).to(dev)
    nv = Fv.shape[1]
    Fw = torch.empty(nv, dtype=torch.float64, device=dev)
    Fe = torch.empty(nv, dtype=torch.float64, device=dev)
    _lib.check(lib.sfxm_fv_fluxes(_scheme_ids[scheme], nv, Fv.data_ptr(), uv.data_ptr(), sigma, _lib.ptr(eig),
                                  Fw.data_ptr(), Fe.data_ptr(), _lib.stream_ptr()))
    return Fw.cpu().numpy(), Fe.cpu().numpy()


def diffusion_fluxes(D, cell_sub, scheme):
    """West and east diffusion fluxes (central differences of D over half the cell distance, flux.pyx:263-327)."""
    _check_stencil(cell_sub, scheme)
    lib, dev = _lib.load(), _device()
    Dv = torch.from_numpy(_values(D, cell_sub)).to(dev)
    xs = torch.from_numpy(np.array([cell.x() for cell in cell_sub], dtype=np.double)).to(dev)
    nv = Dv.shape[1]
    Dw = torch.empty(nv, dtype=torch.float64, device=dev)
    De = torch.empty(nv, dtype=torch.float64, device=dev)
    _lib.check(lib.sfxm_fv_diffusion_fluxes(nv, Dv.data_ptr(), xs.data_ptr(), Dw.data_ptr(), De.data_ptr(), _lib.stream_ptr()))
    return Dw.cpu().numpy(), De.cpu().numpy()
