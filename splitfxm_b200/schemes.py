"""Discretisation scheme names (reference: splitfxm/schemes.py:4-23).

MUSCL / ENO / WENO / QUICK are names only in the reference (no stencil size, no code): Simulation rejects
them (simulation.py:74-75) and so does this package.
"""
from enum import Enum

from . import _lib
from .error import SFXM

FDSchemes = Enum("FDSchemes", "CENTRAL RIGHT_BIAS")
FVSchemes = Enum("FVSchemes", "LAX_FRIEDRICHS CENTRAL UPWIND MUSCL ENO WENO QUICK")

stencil_sizes = {
    FDSchemes.CENTRAL: 3,
    FDSchemes.RIGHT_BIAS: 4,
    FVSchemes.LAX_FRIEDRICHS: 3,
    FVSchemes.CENTRAL: 3,
    FVSchemes.UPWIND: 3,
}

# scheme -> kernel id in the C ABI (include/sfxm_b200.h)
_scheme_ids = {
    FVSchemes.LAX_FRIEDRICHS: _lib.FV_LAX_FRIEDRICHS,
    FVSchemes.CENTRAL: _lib.FV_CENTRAL,
    FVSchemes.UPWIND: _lib.FV_UPWIND,
    FDSchemes.CENTRAL: _lib.FD_CENTRAL,
    FDSchemes.RIGHT_BIAS: _lib.FD_RIGHT_BIAS,
}


def default_scheme(method):
    if method == "FVM":
        return FVSchemes.LAX_FRIEDRICHS
    if method == "FDM":
        return FDSchemes.CENTRAL
    raise SFXM("Invalid numerical method specified")
