"""sfxm_residual_warps (include/sfxm_b200.h): the warp-cooperative residual kernel on a piece of the mesh.  The pieces a
slab-decomposed mesh launches separately - interior warps, then both end warps - must reproduce the one-launch residual
and Euler update bit for bit; models that do not use that kernel are refused."""
import ctypes

import numpy as np
import pytest

from helpers import load, product_simulation

pytestmark = pytest.mark.gpu


def test_pieces_equal_the_whole():
    import torch
    from splitfxm_b200 import _lib
    z = load("sod_cons_nx100_t0p05")
    s, d, m = product_simulation(z)
    eng = s._engine()
    y = s._state(eng, False, None)
    P, _ = eng.problem(False, None)
    lib = eng.lib
    nw = int(lib.sfxm_residual_warp_count(ctypes.byref(P)))
    assert nw == (eng.nx + 29) // 30 and nw > 2
    dt = 1e-4
    r_ref = torch.empty_like(y)
    y_ref = torch.empty_like(y)
    _lib.check(lib.sfxm_residual(ctypes.byref(P), _lib.ptr(eng.x_int), _lib.ptr(eng.params), eng.param_stride, None,
                                 _lib.ptr(y), _lib.ptr(r_ref), _lib.ptr(y_ref), dt, _lib.stream_ptr()))
    for pieces in ([(1, nw - 1), (-1, nw)], [(0, 1), (1, 2), (2, nw)], [(0, nw)]):
        r = torch.full_like(y, float("nan"))
        yo = torch.full_like(y, float("nan"))
        for lo, hi in pieces:
            _lib.check(lib.sfxm_residual_warps(ctypes.byref(P), _lib.ptr(eng.x_int), _lib.ptr(eng.params), eng.param_stride,
                                               _lib.ptr(y), _lib.ptr(r), _lib.ptr(yo), dt, lo, hi, _lib.stream_ptr()))
        torch.cuda.synchronize()
        assert torch.equal(r, r_ref) and torch.equal(yo, y_ref), pieces


def test_other_kernels_are_refused():
    import torch
    from splitfxm_b200 import _lib
    from splitfxm_b200.error import SFXM
    z = load("ad_fdm_central_nx250_split1")
    s, d, m = product_simulation(z)
    eng = s._engine()
    y = s._state(eng, False, None)
    P, _ = eng.problem(False, None)
    out = torch.empty_like(y)
    with pytest.raises(SFXM):
        _lib.check(eng.lib.sfxm_residual_warps(ctypes.byref(P), _lib.ptr(eng.x_int), _lib.ptr(eng.params), eng.param_stride,
                                               _lib.ptr(y), _lib.ptr(out), None, 0.0, 0, 1, _lib.stream_ptr()))
