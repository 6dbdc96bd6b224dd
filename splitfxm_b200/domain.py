"""Domain: the 1-D mesh with ghost cells (reference: splitfxm/domain.py:11-327).

Storage is two contiguous FP64 arrays - positions ``_x[nbl+nx+nbr]`` and values ``_U[nbl+nx+nbr, nv]``
(cell-major, variable-minor: the reference's unsplit vector order, simulation.py:143) - which is what the
CUDA path uploads and downloads in one copy each.  ``Cell`` / ``Boundary`` objects are views of one slot, so
the reference's object-style access (``d.interior()[i].set_value(j, v)``) reads and writes the same memory.
"""
import copy

import numpy as np

from .boundary import Boundary
from .cell import Cell
from .constants import btype
from .error import SFXM


class Domain:
    def __init__(self, cells, left_boundaries, right_boundaries, components, xmin: float = 0.0, xmax: float = 1.0):
        # same argument meaning as domain.py:27-45: left_boundaries[0] is the ghost nearest the interior
        self._nb = {btype.LEFT: int(len(left_boundaries)), btype.RIGHT: int(len(right_boundaries))}
        self._components = components
        self._xmin = xmin
        self._xmax = xmax
        self._adopt([*reversed(left_boundaries), *cells, *right_boundaries], len(cells))

    def _adopt(self, objs, nx):
        """Copy the objects' data into fresh arrays and bind the objects to their slots."""
        nv = len(self._components)
        n = len(objs)
        x = np.empty(n, dtype=np.float64)
        U = np.zeros((n, nv), dtype=np.float64)
        for k, c in enumerate(objs):
            x[k] = c.x()
            vals = np.asarray(c.values(), dtype=np.float64)
            if vals.shape[0] == nv:
                U[k, :] = vals
            elif vals.shape[0] != 0:
                raise SFXM("Cell value length does not match the number of components")
        self._x, self._U = x, U
        self._nx = nx
        for k, c in enumerate(objs):
            c._bind(self, k)
        self._domain = list(objs)

    @classmethod
    def from_arrays(cls, x_all, U_all, nb_left, nb_right, components, xmin=0.0, xmax=1.0):
        """Array-level constructor (no reference equivalent): wraps positions/values incl. ghosts."""
        self = cls.__new__(cls)
        self._nb = {btype.LEFT: int(nb_left), btype.RIGHT: int(nb_right)}
        self._components = components
        self._xmin, self._xmax = xmin, xmax
        self._x = np.array(x_all, dtype=np.float64)
        self._U = np.array(U_all, dtype=np.float64).reshape(self._x.shape[0], len(components))
        self._nx = self._x.shape[0] - nb_left - nb_right
        self._domain = None   # views are created on first object-style access
        return self

    @classmethod
    def from_size(cls, nx: int, nb_left: int, nb_right: int, components, xmin: float = 0.0, xmax: float = 1.0):
        # domain.py:48-99: interior = linspace(xmin, xmax, nx); initial ghost x uses (xmax - xmin) / nx
        nv = len(components)
        dx = (xmax - xmin) / nx
        xl = [xmin - (i + 1) * dx for i in range(nb_left)][::-1]
        xr = [xmax + (i + 1) * dx for i in range(nb_right)]
        for xg in xl:
            if xg > xmax:
                raise SFXM("Inappropriate boundary type given")
        for xg in xr:
            if xg < xmin:
                raise SFXM("Inappropriate boundary type given")
        x_all = np.concatenate([np.array(xl, dtype=np.float64), np.linspace(xmin, xmax, nx),
                                np.array(xr, dtype=np.float64)])
        return cls.from_arrays(x_all, np.zeros((x_all.shape[0], nv)), nb_left, nb_right, components, xmin, xmax)

    # -- views ------------------------------------------------------------------------------------
    def _views(self):
        if self._domain is None:
            nbl, nbr, n = self._nb[btype.LEFT], self._nb[btype.RIGHT], self._x.shape[0]
            objs = []
            for k in range(n):
                if k < nbl:
                    c = Boundary.__new__(Boundary)
                    c._xmin, c._xmax, c._type = self._xmin, self._xmax, btype.LEFT
                elif k >= n - nbr:
                    c = Boundary.__new__(Boundary)
                    c._xmin, c._xmax, c._type = self._xmin, self._xmax, btype.RIGHT
                else:
                    c = Cell.__new__(Cell)
                c.to_delete = False
                c._bind(self, k)
                objs.append(c)
            self._domain = objs
        return self._domain

    def __deepcopy__(self, memo):
        new = Domain.from_arrays(self._x, self._U, self._nb[btype.LEFT], self._nb[btype.RIGHT],
                                 copy.deepcopy(self._components, memo), self._xmin, self._xmax)
        memo[id(self)] = new
        return new

    # -- reference accessors (domain.py:101-274) -------------------------------------------------
    def xmin(self):
        return self._xmin

    def xmax(self):
        return self._xmax

    def ilo(self):
        return self._nb[btype.LEFT]

    def ihi(self):
        return self._nb[btype.LEFT] + self._nx - 1

    def nb(self, dir: btype):
        return self._nb[dir]

    def cells(self, interior=False):
        return self._views() if not interior else self.interior()

    def boundaries(self):
        nb_left, nb_right = self._nb[btype.LEFT], self._nb[btype.RIGHT]
        dom = self._views()
        return dom[:nb_left][::-1], dom[-nb_right:]

    def interior(self):
        nb_left, nb_right = self._nb[btype.LEFT], self._nb[btype.RIGHT]
        return self._views()[nb_left:-nb_right]

    def set_interior(self, cells):
        nb_left, nb_right = self._nb[btype.LEFT], self._nb[btype.RIGHT]
        dom = self._views()
        left, right = dom[:nb_left], dom[-nb_right:]
        objs = [*left, *cells, *right]
        for c in objs:
            c._unbind()
        self._adopt(objs, len(cells))

    def num_components(self):
        return len(self._components)

    def component_index(self, v: str):
        return self._components.index(v)

    def component_name(self, i: int):
        return self._components[i]

    def positions(self, interior=False):
        nb_left, nb_right = self._nb[btype.LEFT], self._nb[btype.RIGHT]
        xs = self._x[nb_left:self._x.shape[0] - nb_right] if interior else self._x
        return xs.tolist()

    def values(self, interior=False):
        nb_left, nb_right = self._nb[btype.LEFT], self._nb[btype.RIGHT]
        U = self._U[nb_left:self._U.shape[0] - nb_right] if interior else self._U
        return np.array(U)

    def listify_interior(self, split, split_locs=None):
        # domain.py:276-312
        nb_left, nb_right = self._nb[btype.LEFT], self._nb[btype.RIGHT]
        interior_values = self._U[nb_left:self._U.shape[0] - nb_right, :]
        if not split:
            return interior_values.flatten()
        if not split_locs:
            raise SFXM("Split locations must be specified when splitting is enabled")
        split_locs = sorted(set(split_locs))
        if split_locs[-1] > interior_values.shape[1]:
            raise SFXM("Split locations exceed the number of columns in the interior values")
        split_segments = np.split(interior_values, split_locs, axis=1)
        return np.concatenate([segment.flatten() for segment in split_segments])

    def update(self, dt, interior_residual_block):
        nb_left, nb_right = self._nb[btype.LEFT], self._nb[btype.RIGHT]
        self._U[nb_left:self._U.shape[0] - nb_right] += dt * np.asarray(interior_residual_block)

    # -- array-level access used by the CUDA host code ---------------------------------------------
    def interior_x(self):
        nb_left, nb_right = self._nb[btype.LEFT], self._nb[btype.RIGHT]
        return self._x[nb_left:self._x.shape[0] - nb_right]

    def interior_U(self):
        nb_left, nb_right = self._nb[btype.LEFT], self._nb[btype.RIGHT]
        return self._U[nb_left:self._U.shape[0] - nb_right]
