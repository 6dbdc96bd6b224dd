"""Boundary (ghost) cell (reference: splitfxm/boundary.py:6-121): a Cell that knows its side and refuses
positions inside the domain."""
import numpy as np

from .cell import Cell
from .constants import btype
from .error import SFXM


class Boundary(Cell):
    __slots__ = ("_xmin", "_xmax", "_type")

    def __init__(self, x, _btype, value=np.array([]), xmin: float = 0.0, xmax: float = 1.0):
        if x < xmin and _btype == btype.RIGHT or x > xmax and _btype == btype.LEFT:
            raise SFXM("Inappropriate boundary type given")
        super().__init__(x, value)
        self._xmin = xmin
        self._xmax = xmax
        self._type = _btype

    def set_x(self, x: float):
        if x >= self._xmin and x <= self._xmax:
            raise SFXM("Boundary cell cannot have interior X-value")
        if (x < self._xmin and self._type == btype.RIGHT) or (x > self._xmax and self._type == btype.LEFT):
            raise SFXM("Inappropriate boundary type given")
        self._x = x
