"""Cell: a position and an nv-vector of values (reference: splitfxm/cell.py:4-136).

A Cell either owns its data (``Cell(x, values)``, as refine.py:267-270 and the reference's tests build
them) or is a *view* of one slot of an array-backed ``Domain`` - then reads and writes go straight to the
domain's arrays, so code that mutates cells (``cell.set_value``, ``cell._x = ...``, ``cell.values()[k] = v``)
keeps working while the hot path sees plain contiguous arrays.
"""
import numpy as np


class Cell:
    __slots__ = ("_own_x", "_own_value", "_dom", "_idx", "to_delete")

    def __init__(self, x=None, value=np.array([])):
        self._own_x = x
        self._own_value = value
        self._dom = None
        self._idx = -1
        self.to_delete = False

    # -- binding to a domain slot -------------------------------------------------------------
    def _bind(self, dom, idx):
        self._dom, self._idx = dom, idx
        self._own_x, self._own_value = None, None

    def _unbind(self):
        if self._dom is not None:
            self._own_x = float(self._dom._x[self._idx])
            self._own_value = np.array(self._dom._U[self._idx])
            self._dom, self._idx = None, -1

    # `_x` / `_value` are touched directly by reference-style code (tests/test_refine.py:133)
    @property
    def _x(self):
        return self._own_x if self._dom is None else float(self._dom._x[self._idx])

    @_x.setter
    def _x(self, x):
        if self._dom is None:
            self._own_x = x
        else:
            self._dom._x[self._idx] = x

    @property
    def _value(self):
        return self._own_value if self._dom is None else self._dom._U[self._idx]

    @_value.setter
    def _value(self, l):
        if self._dom is None:
            self._own_value = l
        else:
            self._dom._U[self._idx, :] = l

    def __eq__(self, other):
        return self._x == other._x

    def __lt__(self, other):
        return self._x < other._x

    __hash__ = None

    def x(self):
        return self._x

    def values(self):
        return self._value

    def value(self, i: int):
        return self._value[i]

    def set_value(self, i: int, val):
        self._value[i] = val

    def set_values(self, l):
        self._value = l

    def update(self, dt, residual):
        v = self._value
        v += dt * residual
