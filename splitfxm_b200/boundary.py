"""Ghost cell of one side of the domain (reference: splitfxm/boundary.py:6-121): a Cell that knows which side it
sits on and refuses positions that are inside the domain or on the wrong side of it."""
import numpy as np

from .cell import Cell
from .constants import btype
from .error import SFXM


def _wrong_side(x, side, xmin, xmax):
    """A right ghost left of the domain, or a left ghost right of it (boundary.py:29-31, 104-107)."""
    return (side is btype.RIGHT and x < xmin) or (side is btype.LEFT and x > xmax)


class Boundary(Cell):
    __slots__ = ("_xmin", "_xmax", "_type")

    def __init__(self, x, _btype, value=np.array([]), xmin: float = 0.0, xmax: float = 1.0):
        if _wrong_side(x, _btype, xmin, xmax):
            raise SFXM("Inappropriate boundary type given")
        super().__init__(x, value)
        self._xmin, self._xmax, self._type = xmin, xmax, _btype

    def set_x(self, x: float):
        if self._xmin <= x <= self._xmax:
            raise SFXM("Boundary cell cannot have interior X-value")
        if _wrong_side(x, self._type, self._xmin, self._xmax):
            raise SFXM("Inappropriate boundary type given")
        self._x = x
