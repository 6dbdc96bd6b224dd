"""Domain: the 1-D mesh with ghost cells (reference: splitfxm/domain.py:11-327).

Storage is two contiguous FP64 arrays - positions ``x[nbl+nx+nbr]`` and values ``U[nbl+nx+nbr, nv]`` (cell-major,
variable-minor: the reference's unsplit vector order, simulation.py:143) - held TWICE: as CUDA tensors (what the
solver calls read and write) and as NumPy arrays (what object-style host code reads and writes), with validity flags.
``Simulation.evolve / steady_state / get_residuals_from_list / jacobian`` and ``Refiner.refine`` leave their result in
the device copy only; the host copy is refreshed on the first host access after that, and a host write access marks
the device copy stale.  A sequence evolve -> refine -> evolve -> steady_state therefore never moves the state over
PCIe.  ``Cell`` / ``Boundary`` objects are views of one slot, so the reference's object-style access
(``d.interior()[i].set_value(j, v)``, even ``cell._x = ...``) keeps working.
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
        self._init_storage()
        self._adopt([*reversed(left_boundaries), *cells, *right_boundaries], len(cells))

    def _init_storage(self):
        self._hx = self._hU = None          # host copy
        self._dx = self._dU = None          # device copy (torch CUDA tensors) or None
        self._host_ok, self._dev_ok = True, False
        self._mesh_rev = 0                  # bumped whenever the positions may have changed
        self._domain = None

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
        self._set_host(x, U, nx)
        for k, c in enumerate(objs):
            c._bind(self, k)
        self._domain = list(objs)

    # -- the two copies ---------------------------------------------------------------------------
    def _set_host(self, x, U, nx):
        """Host arrays become the truth (new mesh)."""
        self._hx, self._hU, self._nx = x, U, int(nx)
        self._host_ok, self._dev_ok = True, False
        self._dx = self._dU = None
        self._mesh_rev += 1
        self._domain = None

    def _host(self, write=False, positions=False):
        """The host arrays, refreshed from the device if stale.  write: the caller may modify them (the device copy
        becomes stale); positions: ... including x (the mesh revision is bumped)."""
        if not self._host_ok:
            xh, Uh = self._dx.cpu().numpy(), self._dU.cpu().numpy()
            if self._hx is not None and self._hx.shape == xh.shape:
                self._hx[:] = xh                    # keep the array objects that Cell views / user code may hold
                self._hU[:, :] = Uh
            else:
                self._hx, self._hU = xh, Uh
                self._domain = None
            self._host_ok = True
        if write:
            self._dev_ok = False
            if positions:
                self._mesh_rev += 1
        return self._hx, self._hU

    def _device(self, device=None):
        """(x, U) as CUDA tensors [ntot], [ntot, nv], uploaded if the device copy is stale."""
        import torch
        if not self._dev_ok:
            dev = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
            self._dx = torch.from_numpy(self._hx).to(dev)
            self._dU = torch.from_numpy(self._hU).to(dev)
            self._dev_ok = True
        return self._dx, self._dU

    def _adopt_device(self, x_all, U_all, nx=None, mesh_changed=False):
        """Device tensors become the truth (result of a solver call / refinement); no copy to the host."""
        nv = len(self._components)
        self._dx, self._dU = x_all, U_all.reshape(x_all.shape[0], nv)
        self._dev_ok, self._host_ok = True, False
        if nx is not None and int(nx) != self._nx:
            self._nx = int(nx)
            self._domain = None           # cell views are rebuilt on demand
            mesh_changed = True
        if mesh_changed:
            self._mesh_rev += 1

    # `_x` / `_U` are touched directly by host code (Cell views, tests): a read may be followed by an in-place write
    # that cannot be intercepted, so every access through these names counts as a write
    @property
    def _x(self):
        return self._host(write=True, positions=True)[0]

    @_x.setter
    def _x(self, x):
        self._host()
        self._set_host(np.asarray(x, dtype=np.float64), self._hU, self._nx)

    @property
    def _U(self):
        return self._host(write=True)[1]

    @_U.setter
    def _U(self, U):
        self._host()
        self._hU = np.asarray(U, dtype=np.float64)
        self._dev_ok = False

    @classmethod
    def from_arrays(cls, x_all, U_all, nb_left, nb_right, components, xmin=0.0, xmax=1.0):
        """Array-level constructor (no reference equivalent): wraps positions/values incl. ghosts."""
        self = cls.__new__(cls)
        self._nb = {btype.LEFT: int(nb_left), btype.RIGHT: int(nb_right)}
        self._components = components
        self._xmin, self._xmax = xmin, xmax
        self._init_storage()
        x = np.array(x_all, dtype=np.float64)
        U = np.array(U_all, dtype=np.float64).reshape(x.shape[0], len(components))
        self._set_host(x, U, x.shape[0] - nb_left - nb_right)
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
        self._host()
        if self._domain is None:
            nbl, nbr, n = self._nb[btype.LEFT], self._nb[btype.RIGHT], self._hx.shape[0]
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
        x, U = self._host()
        new = Domain.from_arrays(x, U, self._nb[btype.LEFT], self._nb[btype.RIGHT],
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
        x = self._host()[0]
        xs = x[nb_left:x.shape[0] - nb_right] if interior else x
        return xs.tolist()

    def values(self, interior=False):
        nb_left, nb_right = self._nb[btype.LEFT], self._nb[btype.RIGHT]
        U = self._host()[1]
        return np.array(U[nb_left:U.shape[0] - nb_right] if interior else U)

    def listify_interior(self, split, split_locs=None):
        # domain.py:276-312
        nb_left, nb_right = self._nb[btype.LEFT], self._nb[btype.RIGHT]
        U = self._host()[1]
        interior_values = U[nb_left:U.shape[0] - nb_right, :]
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
        U = self._host(write=True)[1]
        U[nb_left:U.shape[0] - nb_right] += dt * np.asarray(interior_residual_block)

    # -- array-level access -------------------------------------------------------------------------
    def interior_x(self):
        """Writable host view of the interior positions (marks the device copy stale)."""
        nb_left, nb_right = self._nb[btype.LEFT], self._nb[btype.RIGHT]
        x = self._host(write=True, positions=True)[0]
        return x[nb_left:x.shape[0] - nb_right]

    def interior_U(self):
        """Writable host view of the interior values (marks the device copy stale)."""
        nb_left, nb_right = self._nb[btype.LEFT], self._nb[btype.RIGHT]
        U = self._host(write=True)[1]
        return U[nb_left:U.shape[0] - nb_right]

    def device_arrays(self, device=None):
        """(x [ntot], U [ntot, nv]) as CUDA tensors - the copy the solver calls work on; read-only for callers."""
        return self._device(device)
