"""Discretisation schemes: names, stencil widths, C-ABI kernel ids (reference: splitfxm/schemes.py:4-23).

The reference lists MUSCL / ENO / WENO / QUICK as names only - no stencil size, no code, and Simulation rejects
them (simulation.py:74-75).  They stay names here too and are rejected the same way.
"""
import enum

from . import _lib
from .error import SFXM


class FDSchemes(enum.Enum):
    CENTRAL = 1
    RIGHT_BIAS = 2


class FVSchemes(enum.Enum):
    LAX_FRIEDRICHS = 1
    CENTRAL = 2
    UPWIND = 3
    MUSCL = 4
    ENO = 5
    WENO = 6
    QUICK = 7


# (scheme, cells in the stencil = nb_left + 1 + nb_right, kernel id in include/sfxm_b200.h)
_IMPLEMENTED = (
    (FVSchemes.LAX_FRIEDRICHS, 3, _lib.FV_LAX_FRIEDRICHS),
    (FVSchemes.CENTRAL, 3, _lib.FV_CENTRAL),
    (FVSchemes.UPWIND, 3, _lib.FV_UPWIND),
    (FDSchemes.CENTRAL, 3, _lib.FD_CENTRAL),
    (FDSchemes.RIGHT_BIAS, 4, _lib.FD_RIGHT_BIAS),
)
stencil_sizes = {scheme: width for scheme, width, _ in _IMPLEMENTED}
_scheme_ids = {scheme: kernel for scheme, _, kernel in _IMPLEMENTED}

_DEFAULTS = {"FVM": FVSchemes.LAX_FRIEDRICHS, "FDM": FDSchemes.CENTRAL}


def default_scheme(method):
    """'FVM' -> Lax-Friedrichs, 'FDM' -> central differences (schemes.py:17-23)."""
    try:
        return _DEFAULTS[method]
    except KeyError:
        raise SFXM("Invalid numerical method specified") from None
