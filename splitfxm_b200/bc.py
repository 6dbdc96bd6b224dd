"""Boundary conditions on the host-side Domain view (reference: splitfxm/bc.py:6-240).

``apply_BC`` here serves object-style callers (``Simulation.__init__``, user code between solver calls); inside
``evolve`` / ``steady_state`` / ``jacobian`` / ``get_residuals_from_list`` the ghost cells are computed in the
CUDA kernels themselves (csrc/sfxm_device.cuh: ghost_u_left/right), from the descriptor ``describe_bcs``
builds.
"""
from . import _lib
from .constants import btype, btype_map
from .domain import Domain
from .error import SFXM


def _parse_side(bc_type):
    """-> (C-ABI type id, value) for one side; raises like bc.py:98-159."""
    if bc_type == "periodic":
        return _lib.BC_PERIODIC, 0.0
    if bc_type == "outflow":
        return _lib.BC_OUTFLOW, 0.0
    if isinstance(bc_type, dict) and len(bc_type.keys()) == 1:
        bc_data = list(bc_type.values())[0]
        kind = list(bc_type.keys())[0]
        if not isinstance(bc_data, (float, int)):
            raise SFXM("Incorrect data specified for dictionary-type boundaries")
        if kind == "neumann":
            return _lib.BC_NEUMANN, float(bc_data)
        if kind == "dirichlet":
            return _lib.BC_DIRICHLET, float(bc_data)
        raise SFXM("Incorrect data specified for dictionary-type boundaries")
    raise SFXM("Boundary type not implemented")


def _check_dirs(bc):
    bc_keys = sorted(list(bc.keys()))
    if bc_keys != [btype_map[btype.LEFT], btype_map[btype.RIGHT]]:
        raise SFXM("Incorrect boundary directions specified")


def apply_BC(d: Domain, v: str, bc: dict = {"left": "periodic", "right": "periodic"}):
    """Fill the ghost cells of variable ``v`` (values and positions) - bc.py:6-159."""
    cells = d.cells()
    xmin, xmax = d.xmin(), d.xmax()
    idx = d.component_index(v)
    ilo, ihi = d.ilo(), d.ihi()
    lb, rb = d.boundaries()
    _check_dirs(bc)
    # left, then right (sorted keys)
    kind, val = _parse_side(bc["left"])
    for i, b in enumerate(lb):
        if kind == _lib.BC_PERIODIC:
            b.set_x(xmin - (xmax - cells[ihi - (i + 1)].x()))
            b.set_value(idx, cells[ihi - i].value(idx))
            continue
        b.set_x(xmin - (cells[(i + 1) + ilo].x() - xmin))
        if kind == _lib.BC_OUTFLOW:
            dy = cells[ilo + 1].value(idx) - cells[ilo].value(idx)
            dx = cells[ilo + 1].x() - cells[ilo].x()
            delta_x = cells[ilo - i].x() - cells[ilo - (i + 1)].x()
            b.set_value(idx, cells[ilo - i].value(idx) - (dy / dx) * delta_x)
        elif kind == _lib.BC_NEUMANN:
            dx = cells[ilo - i].x() - cells[ilo - (i + 1)].x()
            b.set_value(idx, cells[ilo - i].value(idx) - val * dx)
        else:
            b.set_value(idx, val)
    kind, val = _parse_side(bc["right"])
    for i, b in enumerate(rb):
        if kind == _lib.BC_PERIODIC:
            b.set_x(xmax + (cells[(i + 1) + ilo].x() - xmin))
            b.set_value(idx, cells[i + ilo].value(idx))
            continue
        b.set_x(xmax + (xmax - cells[ihi - (i + 1)].x()))
        if kind == _lib.BC_OUTFLOW:
            dy = cells[ihi].value(idx) - cells[ihi - 1].value(idx)
            dx = cells[ihi].x() - cells[ihi - 1].x()
            delta_x = cells[ihi + (i + 1)].x() - cells[ihi + i].x()
            b.set_value(idx, cells[ihi + i].value(idx) + (dy / dx) * delta_x)
        elif kind == _lib.BC_NEUMANN:
            dx = cells[ihi + (i + 1)].x() - cells[ihi + i].x()
            b.set_value(idx, cells[ihi + i].value(idx) + val * dx)
        else:
            b.set_value(idx, val)


def get_periodic_bcs(bc_dict: dict, d: Domain):
    """{component index: [sides with a periodic BC]} - bc.py:162-192."""
    periodic_bcs = {}
    for component, bcs in bc_dict.items():
        idx = d.component_index(component)
        sides = [side for side, bc_type in bcs.items() if bc_type == "periodic"]
        if sides:
            periodic_bcs[idx] = sides
    return periodic_bcs


def extend_band(band, dirs: list, i: int, d: Domain):
    """Band indices plus the periodic images - bc.py:195-240."""
    nb_left, nb_right = d.nb(btype.LEFT), d.nb(btype.RIGHT)
    ilo, ihi = d.ilo(), d.ihi()
    addl_points = []
    for dir in dirs:
        if dir == btype_map[btype.RIGHT]:
            overflow = max(0, nb_right - (ihi - i))
            addl_points.extend(list(range(ilo, ilo + overflow)))
        elif dir == btype_map[btype.LEFT]:
            overflow = max(0, nb_left - (i - ilo))
            addl_points.extend(list(range(ihi - overflow + 1, ihi + 1)))
        else:
            raise SFXM("Incorrect boundary direction encountered")
    return list(set(band).union(set(addl_points)))


def describe_bcs(prob, bcs: dict, d: Domain):
    """Fill the BC part of the C-ABI problem descriptor from the reference-style ``bcs`` dict.

    Variables are applied in dict order (simulation.py:69-70, 174-175); the ghost *positions* are shared by all
    variables, so the rule of the LAST variable in that order decides what is stored (bc.py:56-151)."""
    nv = d.num_components()
    nbl, nbr = d.nb(btype.LEFT), d.nb(btype.RIGHT)
    covered = {d.component_index(name) for name in bcs}
    need_ghosts = len(covered) < nv        # a variable without a BC keeps the ghost cells the Domain holds
    if need_ghosts:
        x, U = d._host()                   # read-only host copy (refreshed from the device if stale)
        n = x.shape[0]
    for j in range(nv):
        prob.bc_ltype[j] = _lib.BC_KEEP
        prob.bc_rtype[j] = _lib.BC_KEEP
        for g in range(min(nbl, _lib.MAX_NB)):
            prob.bc_lval[j][g] = float(U[nbl - 1 - g, j]) if need_ghosts else 0.0
        for g in range(min(nbr, _lib.MAX_NB)):
            prob.bc_rval[j][g] = float(U[n - nbr + g, j]) if need_ghosts else 0.0
    prob.gx_left = prob.gx_right = _lib.GX_KEEP
    if not bcs:                            # no BC at all: the ghost positions stay what they are
        x = d._host()[0]
        n = x.shape[0]
        for g in range(min(nbl, _lib.MAX_NB)):
            prob.gx_keep_left[g] = float(x[nbl - 1 - g])
        for g in range(min(nbr, _lib.MAX_NB)):
            prob.gx_keep_right[g] = float(x[n - nbr + g])
    for name, bc in bcs.items():
        j = d.component_index(name)
        _check_dirs(bc)
        lt, lv = _parse_side(bc["left"])
        rt, rv = _parse_side(bc["right"])
        prob.bc_ltype[j], prob.bc_rtype[j] = lt, rt
        for g in range(_lib.MAX_NB):
            prob.bc_lval[j][g] = lv
            prob.bc_rval[j][g] = rv
        prob.gx_left = _lib.GX_PERIODIC if lt == _lib.BC_PERIODIC else _lib.GX_MIRROR
        prob.gx_right = _lib.GX_PERIODIC if rt == _lib.BC_PERIODIC else _lib.GX_MIRROR
