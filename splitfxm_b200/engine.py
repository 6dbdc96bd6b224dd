"""Device engine: one mesh, one model, ``batch`` ensemble members - tensors in HBM plus the C-ABI calls.

Both front ends sit on it: ``Simulation`` (batch = 1, the reference's API) and ``EnsembleSimulation``
(batch = many independent members, no reference equivalent).  PyTorch is used for memory, streams and the
tiny per-member control vectors only; all arithmetic on the state is in csrc/*.cu.
"""
import ctypes
import os

import numpy as np
import torch

from . import _lib
from .error import SFXM


def split_segments(nv, split, split_locs):
    """[(offset, size)] of the variable segments (simulation.py:126-140, 183-196)."""
    if not split:
        return [(0, nv)]
    if split_locs is None:
        raise SFXM("Split location must be specified in this case")
    locs = sorted(set(int(s) for s in split_locs))
    if len(locs) == 0:
        raise SFXM("Split location must be specified in this case")
    if locs[-1] > nv or locs[0] < 0:
        raise SFXM("Split locations must be between 0 and nv-1")
    edges = [0] + locs + [nv]
    return [(edges[k], edges[k + 1] - edges[k]) for k in range(len(edges) - 1)]


def flat_index_table(nx, nv, segs):
    """idx[p, j] = position of unknown (cell p, variable j) in the flat vector."""
    idx = np.empty((nx, nv), dtype=np.int64)
    for (o, s) in segs:
        for j in range(o, o + s):
            idx[:, j] = o * nx + np.arange(nx, dtype=np.int64) * s + (j - o)
    return idx


# Measured on B200 (bench extras, nx = 1e6): with the table the one-mesh kernels are SLOWER (Jacobian 0.153 vs 0.139 ms,
# residual 0.049 vs 0.041 ms - ten more doubles per cell to read), so it stays off for one mesh; SFXM_GEOM_MIN_NX=<nx>
# switches it on from that mesh size for measurements.
GEOM_TABLE_MIN_NX = int(os.environ.get("SFXM_GEOM_MIN_NX", str(1 << 62)))


class Engine:
    def __init__(self, model_id, scheme_id, nv, nbl, nbr, x_int, xmin, xmax, params, batch=1, device=None):
        if not torch.cuda.is_available():
            raise SFXM("splitfxm_b200 needs a CUDA device (there is no CPU fallback)")
        self.lib = _lib.load()
        self.device = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
        self.nv, self.nbl, self.nbr = int(nv), int(nbl), int(nbr)
        self.W = self.nbl + self.nbr + 1
        self.batch = int(batch)
        self.model_id, self.scheme_id = int(model_id), int(scheme_id)
        self.decoupled = self.model_id == _lib.MODEL_AD        # one scalar equation per variable (ModelAD::kDecoupled)
        self.xmin, self.xmax = float(xmin), float(xmax)
        self.set_mesh(x_int)
        self.set_params(params)
        self._bc_filler = None
        self._probs = {}
        self._work = {}

    # -- configuration -------------------------------------------------------------------------------
    def set_mesh(self, x_int):
        """x_int: interior positions, a host array or a CUDA tensor (used in place when contiguous FP64)."""
        if isinstance(x_int, torch.Tensor):
            xt = x_int.to(device=self.device, dtype=torch.float64).contiguous()
        else:
            xt = torch.from_numpy(np.ascontiguousarray(x_int, dtype=np.float64)).to(self.device)
        self.nx = int(xt.shape[0])
        self.n = self.nx * self.nv
        self.x_int = xt
        self._probs = {}
        self._work = {}
        self._geom = None
        self._flat_idx = {}

    def set_params(self, params):
        """params: [n_params] (shared) or [n_params, batch] (per member)."""
        p = np.ascontiguousarray(params, dtype=np.float64)
        last = getattr(self, "_params_host", None)
        if last is not None and last.shape == p.shape and np.array_equal(last, p):
            return
        self._params_host = p.copy()
        if p.ndim == 1:
            self.param_stride = 0
        elif p.ndim == 2 and p.shape[1] == self.batch:
            self.param_stride = self.batch
        else:
            raise SFXM("params must be [n_params] or [n_params, batch]")
        self.n_params = int(p.shape[0])
        self.params = torch.from_numpy(p).to(self.device)

    def set_bcs(self, filler):
        """filler(prob) writes the BC part of a descriptor (bc.describe_bcs)."""
        self._bc_filler = filler
        self._probs = {}
        self._geom = None

    def problem(self, split=False, split_locs=None):
        key = (bool(split), tuple(sorted(set(split_locs))) if (split and split_locs is not None) else None)
        hit = self._probs.get(key)
        if hit is not None:
            return hit
        segs = split_segments(self.nv, split, split_locs)
        P = _lib.Problem()
        P.model, P.scheme = self.model_id, self.scheme_id
        P.nv, P.nbl, P.nbr = self.nv, self.nbl, self.nbr
        P.n_params = self.n_params
        P.nx, P.batch = self.nx, self.batch
        P.xmin, P.xmax = self.xmin, self.xmax
        for k, (o, s) in enumerate(segs):
            for j in range(o, o + s):
                P.var_base[j] = o * self.nx
                P.var_stride[j] = s
                P.var_off[j] = j - o
                P.var_seg[j] = k
        if self._bc_filler is None:
            raise SFXM("boundary conditions not set")
        self._bc_filler(P)
        self._probs[key] = (P, segs)
        if (self.batch >= 32 or (self.decoupled and self.nx >= GEOM_TABLE_MIN_NX)) and self._geom is None:
            # mesh-only data shared by all members: computed once per mesh (the split layout does not enter).  One large
            # mesh of a decoupled model (one thread per unknown) uses the table too: without it every thread derives the
            # scheme's divisors and their reciprocals itself - four IEEE divisions, which made those kernels
            # instruction-bound (0.041 ms for the 56 MB of a 1e6-cell residual)
            self._geom = torch.empty(10 * self.nx, dtype=torch.float64, device=self.device)
            _lib.check(self.lib.sfxm_geometry(ctypes.byref(P), _lib.ptr(self.x_int), _lib.ptr(self._geom), _lib.stream_ptr()))
            _lib.count()
        return P, segs

    # -- flat vector <-> [nx, nv] block, on the device (simulation.py:107-148, domain.py:276-312) -------
    def _flat_index(self, split, split_locs):
        key = (bool(split), tuple(sorted(set(split_locs))) if (split and split_locs is not None) else None)
        hit = self._flat_idx.get(key)
        if hit is None:
            segs = split_segments(self.nv, split, split_locs)
            if len(segs) == 1:
                hit = False                  # unsplit: the flat vector IS the row-major block
            else:
                p = torch.arange(self.nx, dtype=torch.int64, device=self.device)[:, None]
                cols = []
                for (o, s) in segs:
                    for j in range(o, o + s):
                        cols.append(o * self.nx + p * s + (j - o))
                hit = torch.cat(cols, dim=1).reshape(-1)
            self._flat_idx[key] = hit
        return hit

    def block_to_flat(self, Ub, split=False, split_locs=None):
        """[nx, nv] device block -> flat state [n, 1] in the (split-major) vector order; batch 1."""
        idx = self._flat_index(split, split_locs)
        if idx is False:
            return Ub.reshape(self.n, 1).clone()
        y = torch.empty((self.n, 1), dtype=torch.float64, device=self.device)
        y[idx, 0] = Ub.reshape(-1)
        return y

    def flat_to_block(self, y, split=False, split_locs=None):
        idx = self._flat_index(split, split_locs)
        if idx is False:
            return y.reshape(self.nx, self.nv)
        return y.reshape(-1)[idx].reshape(self.nx, self.nv)

    # -- memory ----------------------------------------------------------------------------------------
    def empty_state(self):
        return torch.empty((self.n, self.batch), dtype=torch.float64, device=self.device)

    def to_device(self, l):
        """Host flat vector [n] (batch 1) or [n, batch] -> device [n, batch]."""
        a = np.ascontiguousarray(l, dtype=np.float64)
        if a.ndim == 1:
            if self.batch != 1:
                raise SFXM("state must be [n, batch]")
            a = a.reshape(-1, 1)
        if a.shape != (self.n, self.batch):
            raise SFXM("List length not aligned with interior size")
        return torch.from_numpy(a).to(self.device)

    def jacobian_buffer(self):
        buf = self._work.get("J")
        if buf is None:
            buf = torch.zeros((self.nx, self.W, self.nv, self.nv, self.batch), dtype=torch.float64, device=self.device)
            self._work["J"] = buf
        return buf

    def _scratch(self, name, numel, dtype=torch.float64):
        buf = self._work.get(name)
        if buf is None or buf.numel() < numel:
            buf = torch.empty(max(int(numel), 1), dtype=dtype, device=self.device)
            self._work[name] = buf
        return buf

    # -- C-ABI calls -----------------------------------------------------------------------------------
    def residual(self, y, out=None, split=False, split_locs=None, y_out=None, dt=0.0):
        P, _ = self.problem(split, split_locs)
        if out is None and y_out is None:
            out = torch.empty_like(y)
        _lib.check(self.lib.sfxm_residual(ctypes.byref(P), _lib.ptr(self.x_int), _lib.ptr(self.params), self.param_stride,
                                          _lib.ptr(self._geom), _lib.ptr(y), _lib.ptr(out), _lib.ptr(y_out), float(dt),
                                          _lib.stream_ptr()))
        _lib.count()
        return out

    def euler_step(self, y, y_out, dt, split=False, split_locs=None):
        """y_out = y + dt * f(y) in one launch (simulation.py:455-463)."""
        self.residual(y, None, split, split_locs, y_out=y_out, dt=dt)
        return y_out

    def fill_domain(self, y, split=False, split_locs=None):
        """-> (x_all [ntot], U_all [ntot, nv, batch]) with ghosts, on the device."""
        P, _ = self.problem(split, split_locs)
        ntot = self.nx + self.nbl + self.nbr
        x_all = torch.empty(ntot, dtype=torch.float64, device=self.device)
        U_all = torch.empty((ntot, self.nv, self.batch), dtype=torch.float64, device=self.device)
        _lib.check(self.lib.sfxm_fill_domain(ctypes.byref(P), _lib.ptr(self.x_int), _lib.ptr(y), _lib.ptr(x_all),
                                             _lib.ptr(U_all), _lib.stream_ptr()))
        _lib.count()
        return x_all, U_all

    def jacobian(self, y, split=False, split_locs=None, eps=1e-8, J=None, r0=None, seg_lo=None, seg_hi=None):
        """Block-banded Jacobian; seg_lo/seg_hi restrict the assembly to those variable segments (Split-Newton)."""
        P, segs = self.problem(split, split_locs)
        if J is None:
            J = self.jacobian_buffer()
        if seg_lo is not None and not (seg_lo == 0 and seg_hi == len(segs)):
            _lib.check(self.lib.sfxm_jacobian_segments(ctypes.byref(P), _lib.ptr(self.x_int), _lib.ptr(self.params),
                                                       self.param_stride, _lib.ptr(self._geom), _lib.ptr(y), float(eps),
                                                       int(seg_lo), int(seg_hi), _lib.ptr(J), _lib.ptr(r0), _lib.stream_ptr()))
            _lib.count()
            return J
        _lib.check(self.lib.sfxm_jacobian(ctypes.byref(P), _lib.ptr(self.x_int), _lib.ptr(self.params), self.param_stride,
                                          _lib.ptr(self._geom), _lib.ptr(y), float(eps), _lib.ptr(J), _lib.ptr(r0),
                                          _lib.stream_ptr()))
        _lib.count()
        return J

    def live_entries(self, split=False, split_locs=None):
        """(r, c) pairs of a block that the reference's sparse matrix can hold in this layout, row-major: same
        variable segment (simulation.py:373-392) and, for advection-diffusion, r == c (the other entries are exact
        zeros, which lil_matrix drops)."""
        _, segs = self.problem(split, split_locs)
        seg_of = {}
        for k, (o, s) in enumerate(segs):
            for j in range(o, o + s):
                seg_of[j] = k
        decoupled = self.model_id == _lib.MODEL_AD      # off-diagonal block entries are exact zeros (never stored)
        return [(r, c) for r in range(self.nv) for c in range(self.nv)
                if seg_of[r] == seg_of[c] and (r == c or not decoupled)]

    def pack_jacobian(self, J, split=False, split_locs=None, out=None):
        """-> [nx, W, L, batch] device tensor with the live entries only (sfxm_pack_jacobian)."""
        P, _ = self.problem(split, split_locs)
        L = ctypes.c_int32(0)
        _lib.check(self.lib.sfxm_pack_jacobian(ctypes.byref(P), None, None, ctypes.byref(L), None))
        if out is None:
            out = torch.empty((self.nx, self.W, L.value, self.batch), dtype=torch.float64, device=self.device)
        _lib.check(self.lib.sfxm_pack_jacobian(ctypes.byref(P), _lib.ptr(J), _lib.ptr(out), ctypes.byref(L), _lib.stream_ptr()))
        _lib.count()
        return out

    def solve(self, J, f, s, split=False, split_locs=None, seg_lo=0, seg_hi=None, shift=None, active=None, status=None,
              method=None):
        """(J_seg + shift I) s = -f for segments [seg_lo, seg_hi)."""
        P, segs = self.problem(split, split_locs)
        if seg_hi is None:
            seg_hi = len(segs)
        if self.W == 4:
            method = "band4"          # 4-cell stencil: two-cell super-blocks + PCR
        if method is None:
            # one thread per member (Thomas) needs thousands of members to fill the GPU; below that, one thread per
            # (cell, member) and log2(nx) passes (parallel cyclic reduction)
            method = "thomas" if self.batch >= 2048 else "pcr"
        if method == "band4":
            nbytes = self.lib.sfxm_band4_workspace_bytes(ctypes.byref(P))
            work = self._scratch("band4", nbytes // 8)
            fn = self.lib.sfxm_block_band4_solve_pcr
            _lib.count((seg_hi - seg_lo) * (2 + max((self.nx + 1) // 2 - 1, 1).bit_length()))
        elif method == "thomas":
            nbytes = self.lib.sfxm_solve_workspace_bytes(ctypes.byref(P))
            work = self._scratch("solve", nbytes // 8)
            fn = self.lib.sfxm_block_tridiag_solve
            _lib.count(seg_hi - seg_lo)
        else:
            nbytes = self.lib.sfxm_pcr_workspace_bytes(ctypes.byref(P))
            work = self._scratch("pcr", nbytes // 8)
            fn = self.lib.sfxm_block_tridiag_solve_pcr
            for k in range(seg_lo, seg_hi):
                _lib.count(self._pcr_launches(P, segs[k]))
        _lib.check(fn(ctypes.byref(P), _lib.ptr(J), _lib.ptr(f), _lib.ptr(shift), _lib.ptr(active), int(seg_lo), int(seg_hi),
                      _lib.ptr(s), _lib.ptr(work), _lib.ptr(status), _lib.stream_ptr()))
        return s

    def fused_solve_supported(self, split=False, split_locs=None):
        """True when the Newton linear solve can run with the Jacobian assembled inside the Thomas sweep
        (sfxm_newton_solve_fused): ensembles large enough for one thread per member, built-in 3-cell models."""
        P, _ = self.problem(split, split_locs)
        return self.batch >= 2048 and bool(self.lib.sfxm_newton_fused_supported(ctypes.byref(P)))

    def newton_solve_fused(self, y, s, split=False, split_locs=None, seg=0, eps=1e-8, shift=None, active=None,
                           status=None, r0=None):
        """(J_seg(y) + shift I) s = -f(y) for one segment without materialising J; r0 (optional) receives f(y)."""
        P, _ = self.problem(split, split_locs)
        nbytes = self.lib.sfxm_solve_workspace_bytes(ctypes.byref(P))
        work = self._scratch("solve", nbytes // 8)
        _lib.check(self.lib.sfxm_newton_solve_fused(ctypes.byref(P), _lib.ptr(self.x_int), _lib.ptr(self.params),
                                                    self.param_stride, _lib.ptr(self._geom), _lib.ptr(y), float(eps),
                                                    _lib.ptr(shift), _lib.ptr(active), int(seg), _lib.ptr(s), _lib.ptr(r0),
                                                    _lib.ptr(work), _lib.ptr(status), _lib.stream_ptr()))
        _lib.count()
        return s

    def newton_step_fused(self, y, s, split=False, split_locs=None, seg=0, eps=1e-8, shift=None, active=None,
                          status=None, r0=None):
        """newton_solve_fused with the update folded into the back substitution: y += s in place for the active members;
        -> (sum s^2, sum y_new^2) over the segment, per member (valid for the active members)."""
        P, _ = self.problem(split, split_locs)
        nbytes = self.lib.sfxm_solve_workspace_bytes(ctypes.byref(P))
        work = self._scratch("solve", nbytes // 8)
        out2 = self._scratch("norms_step", 2 * self.batch)[:2 * self.batch]
        _lib.check(self.lib.sfxm_newton_step_fused(ctypes.byref(P), _lib.ptr(self.x_int), _lib.ptr(self.params),
                                                   self.param_stride, _lib.ptr(self._geom), _lib.ptr(y), float(eps),
                                                   _lib.ptr(shift), _lib.ptr(active), int(seg), _lib.ptr(s), _lib.ptr(r0),
                                                   _lib.ptr(work), _lib.ptr(status), _lib.ptr(out2), _lib.stream_ptr()))
        _lib.count()
        return out2[:self.batch], out2[self.batch:]

    def _pcr_launches(self, P, seg, n=None):
        """Kernel launches of one PCR segment solve (csrc/sfxm_pcr.cu): 1 when the rows fit in shared memory
        (pcr_smem_kernel); large meshes: chunk reduce + separator rows + fold around the solve of the nx / 256
        separator system; otherwise init + ceil(log2 n) passes + write-back (+ 1 for the cyclic closure)."""
        import os
        o, S = seg
        n = self.nx if n is None else n
        if n >= 8 * 256 and os.environ.get("SFXM_PCR_CHUNKED", "1")[:1] != "0":
            return 3 + self._pcr_launches(P, seg, (n + 255) // 256)
        cyclic = any(P.bc_ltype[j] == _lib.BC_PERIODIC or P.bc_rtype[j] == _lib.BC_PERIODIC for j in range(o, o + S))
        ne = n - 1 if cyclic else n
        row = 2 * S * S + S * ((1 + S) if cyclic else 1)
        smem = (ne * (row | 1) + S) * 8
        if ne <= (1024 if S <= 2 else 512) and smem <= 200 * 1024 and os.environ.get("SFXM_PCR_SMEM", "1")[:1] != "0":
            return 1
        return 2 + max(ne - 1, 1).bit_length() + (1 if cyclic else 0)

    def sumsq(self, v, lo=0, hi=None, out=None):
        if hi is None:
            hi = self.n
        if out is None:
            out = torch.empty(self.batch, dtype=torch.float64, device=self.device)
        nbytes = self.lib.sfxm_sumsq_workspace_bytes(hi - lo, self.batch)
        work = self._scratch("sumsq", nbytes // 8)
        _lib.check(self.lib.sfxm_sumsq(_lib.ptr(v), int(lo), int(hi), self.batch, _lib.ptr(out), _lib.ptr(work),
                                       _lib.stream_ptr()))
        _lib.count(2)
        return out

    def update_norms(self, y, s, active=None, lo=0, hi=None, update=True, alpha=None):
        """y[lo:hi] += alpha * s[lo:hi] for the active members (if update), and the per-member sums of squares of s and
        of the updated y over [lo, hi) - one pass, two launches (sfxm_update_norms).  -> (sum s^2, sum y^2)."""
        if hi is None:
            hi = self.n
        out2 = torch.empty(2 * self.batch, dtype=torch.float64, device=self.device)
        nbytes = 2 * self.lib.sfxm_sumsq_workspace_bytes(hi - lo, self.batch)
        work = self._scratch("norms2", nbytes // 8)
        _lib.check(self.lib.sfxm_update_norms(_lib.ptr(y), _lib.ptr(s), _lib.ptr(alpha), _lib.ptr(active), 1 if update else 0,
                                              int(lo), int(hi), self.batch, _lib.ptr(out2), _lib.ptr(work),
                                              _lib.stream_ptr()))
        _lib.count(2)
        return out2[:self.batch], out2[self.batch:]

    def axpy(self, y, s, alpha=None, active=None, lo=0, hi=None):
        if hi is None:
            hi = self.n
        _lib.check(self.lib.sfxm_axpy_masked(_lib.ptr(y), _lib.ptr(s), _lib.ptr(alpha), _lib.ptr(active), int(lo), int(hi),
                                             self.batch, _lib.stream_ptr()))
        _lib.count()
        return y

    # -- export ---------------------------------------------------------------------------------------
    def jacobian_coo(self, J_host, split=False, split_locs=None):
        """Banded J (host array [nx, W, nv, nv], one member) -> (rows, cols, vals) in the reference's index
        convention (simulation.py:357-392): exact zeros dropped (lil_matrix never stores them), split mode keeps
        same-segment entries only."""
        _, segs = self.problem(split, split_locs)
        nx, nv, W, nbl = self.nx, self.nv, self.W, self.nbl
        idx = flat_index_table(nx, nv, segs)
        seg_of = np.empty(nv, dtype=np.int64)
        for k, (o, s) in enumerate(segs):
            seg_of[o:o + s] = k
        i = np.arange(nx)[:, None, None, None]
        d = np.arange(W)[None, :, None, None]
        r = np.arange(nv)[None, None, :, None]
        c = np.arange(nv)[None, None, None, :]
        pcol = (i + d - nbl) % nx
        rows = np.broadcast_to(idx[i, r], J_host.shape)
        cols = idx[pcol, c]
        cols = np.broadcast_to(cols, J_host.shape)
        keep = (J_host != 0.0) & np.broadcast_to(seg_of[r] == seg_of[c], J_host.shape)
        return rows[keep], cols[keep], J_host[keep]
