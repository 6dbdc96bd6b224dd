"""Shared host side of the per-cell equation objects: evaluate the user's Python callables on the stencil cells (they
can only run on the host - the reference's wrappers do the same, flux.pyx:102-103, derivatives.pyx:214-215) and upload
the values; the arithmetic is one kernel launch (csrc/sfxm_stencil_api.cu)."""
import numpy as np
import torch

from ..error import SFXM


def device():
    if not torch.cuda.is_available():
        raise SFXM("splitfxm_b200 needs a CUDA device (there is no CPU fallback)")
    return torch.device("cuda", torch.cuda.current_device())


def stencil_values(fn, cell_sub, dev):
    return torch.from_numpy(np.ascontiguousarray(np.array([fn(cell.values()) for cell in cell_sub], dtype=np.double))).to(dev)


def stencil_x(cell_sub, dev):
    return torch.from_numpy(np.array([cell.x() for cell in cell_sub], dtype=np.double)).to(dev)


def source_values(S, cell_sub, nv, dev):
    """S receives the list of cells (fv_transport.py:58) and may return a scalar, [0.0] or an nv-vector (the tutorials'
    Euler models return zeros_like of the cell list): broadcast to nv."""
    s = np.asarray(S(cell_sub), dtype=np.float64).reshape(-1)
    if s.shape[0] not in (1, nv):
        raise SFXM("source term must be a scalar or one value per variable")
    return torch.from_numpy(np.ascontiguousarray(np.broadcast_to(s, (nv,)))).to(dev)
