"""NumPy restatement of the reference's residual / boundary-fill / finite-difference-Jacobian path.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  All citations are relative to the reference tree
(gpavanb1/SplitFXM v0.5.2).  Arrays: ``x_int[nx]`` interior positions, ``U_int[nx, nv]`` interior values
(cell-major, variable-minor - the reference's unsplit flat vector, simulation.py:143).

Operation order is kept exactly as in the compiled Cython / NumPy reference so that, for models whose
F/D are a single multiply (advection-diffusion), results are bit-identical.
"""
from __future__ import annotations

import json

import numpy as np

EPS_FD = 1e-8  # simulation.py:270 (epsilon default)

# ------------------------------------------------------------------------------------------------
# Boundary conditions (bc.py:6-159)
# ------------------------------------------------------------------------------------------------
PERIODIC, OUTFLOW, NEUMANN, DIRICHLET, NONE = 0, 1, 2, 3, 4


class OracleError(Exception):
    """Mirrors splitfxm.error.SFXM (error.py:1)."""


def parse_side(spec):
    """One side of one variable's BC -> (type, value).  bc.py:47-159."""
    if spec == "periodic":
        return PERIODIC, 0.0
    if spec == "outflow":
        return OUTFLOW, 0.0
    if isinstance(spec, dict) and len(spec) == 1:
        (k, v), = spec.items()
        if not isinstance(v, (float, int)):
            raise OracleError("Incorrect data specified for dictionary-type boundaries")
        if k == "neumann":
            return NEUMANN, float(v)
        if k == "dirichlet":
            return DIRICHLET, float(v)
        raise OracleError("Incorrect data specified for dictionary-type boundaries")
    raise OracleError("Boundary type not implemented")


def parse_bcs(bcs, comps):
    """bcs dict (insertion-ordered, as Simulation.__init__ iterates it, simulation.py:69-70) ->
    list of (var_index, (ltype, lval), (rtype, rval)) in application order."""
    if isinstance(bcs, str):
        bcs = json.loads(bcs)
    out = []
    comps = list(comps)
    for name, bc in bcs.items():
        if sorted(bc.keys()) != ["left", "right"]:
            raise OracleError("Incorrect boundary directions specified")
        out.append((comps.index(name), parse_side(bc["left"]), parse_side(bc["right"])))
    return out


def fill_ghosts(x_int, U_int, nbl, nbr, bc_list, xmin, xmax, ghost_x0=None, ghost_U0=None):
    """apply_BC for every variable in order (bc.py:6-159) on a fresh copy.

    Returns (x_all[nbl+nx+nbr], U_all[nbl+nx+nbr, nv]).  Ghost k (0 = nearest the interior) sits at index
    ilo-1-k on the left and ihi+1+k on the right (domain.py:167-177).  Ghost *positions* are shared by all
    variables, so the stored ghost x is whatever the last variable in ``bc_list`` wrote (bc.py:56,64,73,
    87,114,127,141,151); the *values* of outflow/neumann ghosts use the ghost x their own variable wrote.
    ``ghost_x0``/``ghost_U0`` seed ghosts of variables/sides no BC touches (domain.py:87-97).
    """
    nx, nv = U_int.shape
    n = nbl + nx + nbr
    x = np.empty(n)
    U = np.zeros((n, nv))
    ilo, ihi = nbl, nbl + nx - 1
    x[ilo:ihi + 1] = x_int
    U[ilo:ihi + 1] = U_int
    if ghost_x0 is not None:
        x[:ilo] = ghost_x0[:ilo]
        x[ihi + 1:] = ghost_x0[ihi + 1:]
    else:
        # Domain.from_size initial ghost positions (domain.py:87-97)
        dx0 = (xmax - xmin) / nx
        for i in range(nbl):
            x[ilo - 1 - i] = xmin - (i + 1) * dx0
        for i in range(nbr):
            x[ihi + 1 + i] = xmax + (i + 1) * dx0
    if ghost_U0 is not None:
        U[:ilo] = ghost_U0[:ilo]
        U[ihi + 1:] = ghost_U0[ihi + 1:]
    for (j, (lt, lv), (rt, rv)) in bc_list:
        # ---- left (bc.py:50-57, 67-80, 107-119, 134-143)
        for i in range(nbl):
            g = ilo - 1 - i
            if lt == PERIODIC:
                x[g] = xmin - (xmax - x[ihi - (i + 1)])
                U[g, j] = U[ihi - i, j]
            else:
                x[g] = xmin - (x[(i + 1) + ilo] - xmin)
                if lt == OUTFLOW:
                    dy = U[ilo + 1, j] - U[ilo, j]
                    dx = x[ilo + 1] - x[ilo]
                    delta_x = x[ilo - i] - x[ilo - (i + 1)]
                    U[g, j] = U[ilo - i, j] - (dy / dx) * delta_x
                elif lt == NEUMANN:
                    dx = x[ilo - i] - x[ilo - (i + 1)]
                    U[g, j] = U[ilo - i, j] - lv * dx
                elif lt == DIRICHLET:
                    U[g, j] = lv
        # ---- right (bc.py:59-65, 82-94, 121-132, 144-153)
        for i in range(nbr):
            g = ihi + 1 + i
            if rt == PERIODIC:
                x[g] = xmax + (x[(i + 1) + ilo] - xmin)
                U[g, j] = U[i + ilo, j]
            else:
                x[g] = xmax + (xmax - x[ihi - (i + 1)])
                if rt == OUTFLOW:
                    dy = U[ihi, j] - U[ihi - 1, j]
                    dx = x[ihi] - x[ihi - 1]
                    delta_x = x[ihi + (i + 1)] - x[ihi + i]
                    U[g, j] = U[ihi + i, j] + (dy / dx) * delta_x
                elif rt == NEUMANN:
                    dx = x[ihi + (i + 1)] - x[ihi + i]
                    U[g, j] = U[ihi + i, j] + rv * dx
                elif rt == DIRICHLET:
                    U[g, j] = rv
    return x, U


# ------------------------------------------------------------------------------------------------
# Models (models/advection_diffusion.py:28-31, tutorials/*/euler1d.py)
# ------------------------------------------------------------------------------------------------
class AdvectionDiffusion:
    """F = c u, D = nu u, S = 0, dFdU = diag(c)  (models/advection_diffusion.py:28-31)."""
    name = "ad"
    diagonal_dfdu = True

    def __init__(self, c, nu):
        self.c, self.nu = float(c), float(nu)

    def F(self, U):
        return self.c * U

    def D(self, U):
        return self.nu * U

    def sigma(self, U):
        # max|eigvals(diag(c))| = |c| exactly (flux.pyx:98-99)
        return np.full(U.shape[:-1], abs(self.c))

    def eig_diag(self, U):
        return np.full(U.shape, self.c)


class Euler1D:
    """Primitive (rho, u, p); dFdU is dF/dW (tutorials/sod_shock/euler1d.py:27-50)."""
    name = "euler_prim"
    diagonal_dfdu = False

    def __init__(self, gamma=1.4):
        self.gamma = float(gamma)

    def F(self, U):
        g = self.gamma
        rho, u, p = U[..., 0], U[..., 1], U[..., 2]
        E = p / (g - 1) + 0.5 * rho * u ** 2
        return np.stack([rho * u, rho * u ** 2 + p, u * (E + p)], axis=-1)

    def D(self, U):
        return np.zeros_like(U)

    def dFdU(self, U):
        g = self.gamma
        rho, u, p = U[..., 0], U[..., 1], U[..., 2]
        z = np.zeros_like(rho)
        o = np.ones_like(rho)
        return np.stack([
            np.stack([u, rho, z], -1),
            np.stack([u ** 2, 2 * rho * u, o], -1),
            np.stack([0.5 * (u ** 3.), 1.5 * rho * (u ** 2.) + g * p / (g - 1.), g * u / (g - 1.)], -1)], -2)

    def sigma(self, U):
        # LAPACK per cell in the reference (flux.pyx:98-99)
        return np.max(np.abs(np.linalg.eigvals(self.dFdU(U))), axis=-1)


class Euler1DConservative:
    """Conservative (rho, rho u, E) (tutorials/conservative_sod_shock/euler1d.py:28-63)."""
    name = "euler_cons"
    diagonal_dfdu = False

    def __init__(self, gamma=1.4):
        self.gamma = float(gamma)

    def F(self, U):
        g = self.gamma
        U1, U2, U3 = U[..., 0], U[..., 1], U[..., 2]
        F1 = U2
        F2 = 0.5 * (3 - g) * U2 ** 2 / U1 + (g - 1) * U3
        F3 = g * U2 * U3 / U1 - 0.5 * (g - 1) * (U2 ** 3 / U1 ** 2)
        return np.stack([F1, F2, F3], axis=-1)

    def D(self, U):
        return np.zeros_like(U)

    def dFdU(self, U):
        g = self.gamma
        U1, U2, U3 = U[..., 0], U[..., 1], U[..., 2]
        z = np.zeros_like(U1)
        o = np.ones_like(U1)
        D4 = -0.5 * (3 - g) * (U2 ** 2 / U1 ** 2)
        D5 = (3 - g) * U2 / U1
        D6 = (g - 1) * o
        D7 = -g * U2 * U3 / U1 ** 2 + (g - 1) * (U2 ** 3 / U1 ** 3)
        D8 = g * U3 / U1 - 1.5 * (g - 1) * (U2 ** 2 / U1 ** 2)
        D9 = g * U2 / U1
        return np.stack([np.stack([z, o, z], -1), np.stack([D4, D5, D6], -1), np.stack([D7, D8, D9], -1)], -2)

    def sigma(self, U):
        return np.max(np.abs(np.linalg.eigvals(self.dFdU(U))), axis=-1)


def make_model(name, **kw):
    if name == "ad":
        return AdvectionDiffusion(kw["c"], kw["nu"])
    if name == "euler_prim":
        return Euler1D(kw.get("gamma", 1.4))
    if name == "euler_cons":
        return Euler1DConservative(kw.get("gamma", 1.4))
    raise OracleError("unknown model " + name)


# ------------------------------------------------------------------------------------------------
# Stencil kernels (flux.pyx, derivatives.pyx, equations/*.py) - vectorised over cells
# ------------------------------------------------------------------------------------------------
FV_SCHEMES = ("LAX_FRIEDRICHS", "CENTRAL", "UPWIND")
FD_SCHEMES = ("CENTRAL", "RIGHT_BIAS")


def stencil_size(method, scheme):
    """schemes.py:8-14."""
    if method == "FDM":
        return {"CENTRAL": 3, "RIGHT_BIAS": 4}[scheme]
    return {"LAX_FRIEDRICHS": 3, "CENTRAL": 3, "UPWIND": 3}[scheme]


def fv_fluxes(model, scheme, ul, uc, ur):
    """fluxes (flux.pyx:42-105): west / east face fluxes of every cell, [n, nv] each, for LAX_FRIEDRICHS (flux.pyx:129-169),
    CENTRAL (flux.pyx:229-257) and UPWIND with a diagonal dFdU (flux.pyx:174-224)."""
    Fl, Fc, Fr = model.F(ul), model.F(uc), model.F(ur)
    if scheme == "LAX_FRIEDRICHS":
        sigma = model.sigma(uc)[:, None]                  # flux.pyx:98-99
        Fw = 0.5 * (Fl + Fc) - 0.5 * sigma * (uc - ul)    # flux.pyx:154-160
        Fe = 0.5 * (Fc + Fr) - 0.5 * sigma * (ur - uc)    # flux.pyx:163-167
    elif scheme == "CENTRAL":
        Fw = 0.5 * (Fl + Fc)                              # flux.pyx:253-255
        Fe = 0.5 * (Fc + Fr)
    elif scheme == "UPWIND":
        if not model.diagonal_dfdu:
            raise OracleError("UPWIND is reproducible for diagonal dFdU only (SURVEY.md 7.3)")
        eig = model.eig_diag(uc)                          # flux.pyx:114, R = I
        pos = eig > 0
        Fw = np.where(pos, Fl, Fc)                        # flux.pyx:211-218
        Fe = np.where(pos, Fc, Fr)
    else:
        raise OracleError("Unsupported scheme")
    return Fw, Fe


def fv_diffusion_fluxes(model, xl, xc, xr, ul, uc, ur):
    """diffusion_fluxes + diff_central (flux.pyx:263-327): note the 0.5 in the face spacings (flux.pyx:285, 294)."""
    Dl, Dc, Dr = model.D(ul), model.D(uc), model.D(ur)
    hw = (0.5 * (xc - xl))[:, None]                       # flux.pyx:285
    he = (0.5 * (xr - xc))[:, None]                       # flux.pyx:294
    return (Dc - Dl) / hw, (Dr - Dc) / he


def fd_dx(scheme, xs, Fs):
    """Dx / dx (derivatives.pyx:75-90, 196-216): first derivative of the slot values Fs[k] ([n, nv]) at positions xs[k]."""
    if scheme == "CENTRAL":
        return (Fs[2] - Fs[0]) / (xs[2] - xs[0])[:, None]                  # derivatives.pyx:75-81
    if scheme == "RIGHT_BIAS":
        h = (xs[1] - xs[0])[:, None]
        return (-Fs[0] - Fs[1] + Fs[2] + Fs[3]) / (4 * h)                  # derivatives.pyx:83-90
    raise OracleError("Unsupported scheme")


def fd_d2x(scheme, xs, Ds):
    """D2x / d2x (derivatives.pyx:96-121, 240-260): second derivative."""
    if scheme == "CENTRAL":
        dxw = (xs[1] - xs[0])[:, None]
        dxe = (xs[2] - xs[1])[:, None]
        Dw = (Ds[1] - Ds[0]) / dxw                                         # derivatives.pyx:96-112
        De = (Ds[2] - Ds[1]) / dxe
        return (De - Dw) / ((dxw + dxe) / 2)
    if scheme == "RIGHT_BIAS":
        h = (xs[1] - xs[0])[:, None]
        return (Ds[0] - Ds[1] - Ds[2] + Ds[3]) / (2 * h ** 2)              # derivatives.pyx:114-121 (first spacing only)
    raise OracleError("Unsupported scheme")


def residual_block(model, method, scheme, x_all, U_all, nbl, nbr):
    """System.residuals (system.py:28-70) for every interior cell at once -> [nx, nv]."""
    n = x_all.shape[0]
    nx = n - nbl - nbr
    if stencil_size(method, scheme) != nbl + nbr + 1:
        raise OracleError("Stencil size does not match boundary conditions")
    S = nbl + nbr + 1
    xs = [x_all[k:k + nx] for k in range(S)]             # xs[k][i] = x of stencil slot k for cell i
    Us = [U_all[k:k + nx] for k in range(S)]
    if method == "FVM":
        # fv_transport.py:49-58
        xl, xc, xr = xs
        ul, uc, ur = Us
        dxw = xc - xl
        dxe = xr - xc
        dx = 0.5 * (dxw + dxe)
        Fw, Fe = fv_fluxes(model, scheme, ul, uc, ur)
        DFw, DFe = fv_diffusion_fluxes(model, xl, xc, xr, ul, uc, ur)
        inv = (1 / dx)[:, None]
        return (0.0 - inv * (Fe - Fw)) + inv * (DFe - DFw)
    # FDM: fd_transport.py:46-50, derivatives.pyx:75-121
    if scheme in FD_SCHEMES:
        dF = fd_dx(scheme, xs, [model.F(u) for u in Us])
        d2 = fd_d2x(scheme, xs, [model.D(u) for u in Us])
        return (0.0 - dF) + d2
    raise OracleError("Unsupported scheme")


# ------------------------------------------------------------------------------------------------
# Split permutation (simulation.py:107-148, 192-200; domain.py:276-312)
# ------------------------------------------------------------------------------------------------
def segments(nv, split, split_locs):
    """-> list of (offset, size) of the variable segments."""
    if not split:
        return [(0, nv)]
    if split_locs is None or len(split_locs) == 0:
        raise OracleError("Split locations must be specified in this case")
    locs = sorted(set(int(s) for s in split_locs))
    if locs[-1] > nv or locs[0] < 0:
        raise OracleError("Split locations must be between 0 and nv-1")
    edges = [0] + locs + [nv]
    return [(edges[k], edges[k + 1] - edges[k]) for k in range(len(edges) - 1)]


def flat_index(nx, nv, split, split_locs):
    """idx[p, j] = position of unknown (cell p, variable j) in the flat vector."""
    idx = np.empty((nx, nv), dtype=np.int64)
    for (o, s) in segments(nv, split, split_locs):
        for j in range(o, o + s):
            idx[:, j] = o * nx + np.arange(nx) * s + (j - o)
    return idx


def to_block(l, nx, nv, split, split_locs):
    """initialize_from_list (simulation.py:107-148): flat -> [nx, nv]."""
    l = np.asarray(l, dtype=np.float64)
    return l[flat_index(nx, nv, split, split_locs)]


def to_flat(block, split, split_locs):
    """listify_interior (domain.py:276-312) / residual re-ordering (simulation.py:192-200)."""
    nx, nv = block.shape
    out = np.empty(nx * nv)
    out[flat_index(nx, nv, split, split_locs)] = block
    return out


# ------------------------------------------------------------------------------------------------
# Problem = everything Simulation holds that the residual depends on
# ------------------------------------------------------------------------------------------------
class Problem:
    def __init__(self, model, method, scheme, x_int, nbl, nbr, bcs, comps, xmin=0.0, xmax=1.0,
                 ghost_x0=None):
        self.model, self.method, self.scheme = model, method, scheme
        self.x_int = np.asarray(x_int, dtype=np.float64)
        self.nx = self.x_int.shape[0]
        self.nbl, self.nbr = int(nbl), int(nbr)
        self.comps = list(comps)
        self.nv = len(self.comps)
        self.bc_list = parse_bcs(bcs, self.comps)
        self.xmin, self.xmax = float(xmin), float(xmax)
        self.ghost_x0 = ghost_x0

    # simulation.py:150-202
    def residual(self, l, split=False, split_locs=None):
        Ub = to_block(l, self.nx, self.nv, split, split_locs)
        return to_flat(self.residual_of_block(Ub), split, split_locs)

    def residual_of_block(self, Ub):
        x_all, U_all = fill_ghosts(self.x_int, Ub, self.nbl, self.nbr, self.bc_list, self.xmin, self.xmax,
                                   self.ghost_x0)
        return residual_block(self.model, self.method, self.scheme, x_all, U_all, self.nbl, self.nbr)

    def ghosts(self, Ub):
        return fill_ghosts(self.x_int, Ub, self.nbl, self.nbr, self.bc_list, self.xmin, self.xmax, self.ghost_x0)

    # simulation.py:270-394 - whole-vector form, one unknown perturbed at a time (SURVEY.md 9.2:
    # identical values AND zero pattern to the reference's per-cell loops)
    def jacobian_dense_unsplit(self, Ub, eps=EPS_FD):
        nx, nv = self.nx, self.nv
        r0 = self.residual_of_block(Ub)
        J = np.zeros((nx * nv, nx * nv))
        for p in range(nx):
            for j in range(nv):
                Up = Ub.copy()
                Up[p, j] = Up[p, j] + eps
                J[:, p * nv + j] = ((self.residual_of_block(Up) - r0) / eps).ravel()
        return J

    def band_columns(self):
        """Cyclic column-cell offsets d = -nbl..nbr of the block band."""
        return list(range(-self.nbl, self.nbr + 1))

    def jacobian_banded(self, Ub, eps=EPS_FD):
        """Block-banded Jacobian Jb[i, d, r, c] = d r_(i,r) / d U_((i+d) mod nx, c), d = -nbl..nbr, built by
        colour passes (cells p == k mod ncol perturbed together, the last cells one at a time so cyclic
        neighbours never share a colour).  Same entries as jacobian_dense_unsplit - see test_oracle_golden."""
        nx, nv, nbl, nbr = self.nx, self.nv, self.nbl, self.nbr
        W = nbl + nbr + 1
        if nx < W:
            raise OracleError("mesh smaller than the stencil")
        r0 = self.residual_of_block(Ub)
        Jb = np.zeros((nx, W, nv, nv))
        # colour classes over cells
        tail = min(nx, 2 * W)
        main = nx - tail
        classes = [np.arange(k, main, W) for k in range(W) if k < main]
        classes += [np.array([p]) for p in range(main, nx)]
        rows_of = np.arange(nx)
        for cls in classes:
            if cls.size == 0:
                continue
            for c in range(nv):
                Up = Ub.copy()
                Up[cls, c] = Up[cls, c] + eps
                dr = (self.residual_of_block(Up) - r0) / eps       # [nx, nv]
                for d in range(-nbl, nbr + 1):
                    rows = (cls - d) % nx                            # rows i with (i+d) mod nx == p
                    Jb[rows, d + nbl, :, c] = dr[rows, :]
        return Jb

    def banded_to_coo(self, Jb, split=False, split_locs=None):
        """Export in the reference's (row, col, val) convention (simulation.py:354-392): exact zeros
        dropped (lil_matrix never stores them), split mode keeps same-segment entries only."""
        nx, nv, nbl = self.nx, self.nv, self.nbl
        idx = flat_index(nx, nv, split, split_locs)
        seg_of = np.zeros(nv, dtype=np.int64)
        for k, (o, s) in enumerate(segments(nv, split, split_locs)):
            seg_of[o:o + s] = k
        rows, cols, vals = [], [], []
        W = Jb.shape[1]
        acc = {}
        for d in range(W):
            pcol = (np.arange(nx) + d - nbl) % nx
            for r in range(nv):
                for c in range(nv):
                    if seg_of[r] != seg_of[c]:
                        continue
                    v = Jb[:, d, r, c]
                    nz = np.nonzero(v)[0]
                    for i in nz:
                        key = (int(idx[i, r]), int(idx[pcol[i], c]))
                        acc[key] = acc.get(key, 0.0) + v[i]
        keys = sorted(acc.keys())
        rows = np.array([k[0] for k in keys], dtype=np.int64)
        cols = np.array([k[1] for k in keys], dtype=np.int64)
        vals = np.array([acc[k] for k in keys])
        return rows, cols, vals

    def jacobian_coo(self, l, split=False, split_locs=None, eps=EPS_FD):
        Ub = to_block(l, self.nx, self.nv, split, split_locs)
        return self.banded_to_coo(self.jacobian_banded(Ub, eps), split, split_locs)

    def jacobian_matrix(self, l, split=False, split_locs=None, eps=EPS_FD):
        """scipy CSR of Simulation.jacobian(l, split, split_locs)."""
        from scipy.sparse import csr_matrix
        r, c, v = self.jacobian_coo(l, split, split_locs, eps)
        n = self.nx * self.nv
        return csr_matrix((v, (r, c)), shape=(n, n))


# ------------------------------------------------------------------------------------------------
# Mesh + initial conditions (domain.py:48-99, initialize.py:7-59)
# ------------------------------------------------------------------------------------------------
def from_size(nx, nbl, nbr, xmin=0.0, xmax=1.0):
    x_int = np.linspace(xmin, xmax, nx)
    dx = (xmax - xmin) / nx
    xl = [xmin - (i + 1) * dx for i in range(nbl)][::-1]
    xr = [xmax + (i + 1) * dx for i in range(nbr)]
    return np.concatenate([xl, x_int, xr])


def initial_condition(x, kind, xmin=0.0, xmax=1.0):
    """initialize.py:36-59 on an array of interior positions."""
    width = xmax - xmin
    x = np.asarray(x, dtype=np.float64)
    if kind == "tophat":
        return np.where((x >= xmin + (0.333 * width)) & (x <= xmin + (0.666 * width)), 1.0, 0.0)
    if kind == "sine":
        inside = (x >= xmin + (0.333 * width)) & (x <= xmin + (0.666 * width))
        return np.where(inside, 1.0 + 0.5 * np.sin(2.0 * np.pi * (x - xmin - 0.333 * width) / (0.333 * width)), 1.0)
    if kind == "rarefaction":
        return np.where(x > xmin + (0.5 * width), 2.0, 1.0)
    if kind == "gaussian":
        return np.exp(-200 * (x - xmin - 0.25 * width) ** 2.0)
    raise NotImplementedError
