"""Adaptive mesh refinement front end (reference: splitfxm/refine.py:10-312, a port of Cantera's refine.cpp).

The criteria bookkeeping and the stdout report live here; marking, the adjacent-deletion rule, the scans and
the compaction into the new mesh run on the device (csrc/sfxm_amr.cu) and reproduce the reference mesh exactly.
"""
import numpy as np

from . import _lib
from .constants import btype
from .error import SFXM


class Refiner:
    def __init__(self):
        # defaults: refine.py:23-36 (borrowed from Cantera)
        self._ratio = 10.0
        self._slope = 0.8
        self._curve = 0.8
        self._prune = -0.001     # a negative prune factor disables pruning
        self._npmax = 1000
        self._min_range = 0.01
        self._min_grid = 1e-10

    def set_criteria(self, ratio, slope, curve, prune):
        # refine.py:38-76
        if ratio < 2.0:
            raise SFXM(f"ratio must be greater than 2.0 ({ratio} was specified).")
        elif slope < 0.0 or slope > 1.0:
            raise SFXM(f"slope must be between 0.0 and 1.0 ({slope} was specified).")
        elif curve < 0.0 or curve > 1.0:
            raise SFXM(f"curve must be between 0.0 and 1.0 ({curve} was specified).")
        elif prune > curve or prune > slope:
            raise SFXM(f"prune must be less than 'curve' and 'slope' ({prune} was specified).")
        self._ratio = ratio
        self._slope = slope
        self._curve = curve
        self._prune = prune

    def set_max_points(self, npmax):
        self._npmax = npmax

    # ------------------------------------------------------------------------------------------------
    def _refine_device(self, xd, Ud, want_flags=True, pad=(0, 0)):
        """Device refinement of one mesh: xd [n], Ud [n, nv] CUDA tensors -> (x_buf, U_buf, n_new, flags) on the device.

        Three launches (sfxm_refine_fused: statistics, marking, single-pass scan + compaction).  The new mesh is written
        into buffers of capacity pad[0] + (2n - 1) + pad[1] cells starting at cell pad[0] (room for the caller's ghost
        cells); the only host round trip is the new size n_new.  flags = dict(loc, keep, gridpt, comp): device byte/int
        arrays for the report, or None."""
        import torch
        lib = _lib.load()
        n, nv = Ud.shape
        if n > self._npmax:
            raise SFXM("Exceeded maximum number of points")        # refine.py:124-125
        dev = xd.device
        work = torch.empty(lib.sfxm_refine_fused_workspace_bytes(n), dtype=torch.uint8, device=dev)
        counts = torch.zeros(3, dtype=torch.int64, device=dev)
        cap = pad[0] + 2 * n - 1 + pad[1]
        x_buf = torch.empty(cap, dtype=torch.float64, device=dev)
        U_buf = torch.empty((cap, nv), dtype=torch.float64, device=dev)
        st = _lib.stream_ptr()
        _lib.check(lib.sfxm_refine_fused(xd.data_ptr(), Ud.data_ptr(), n, nv, float(self._ratio), float(self._slope),
                                         float(self._curve), float(self._prune), float(self._min_range),
                                         float(self._min_grid), work.data_ptr(), x_buf[pad[0]:].data_ptr(),
                                         U_buf[pad[0]:].data_ptr(), counts.data_ptr(), st))
        _lib.count(_lib.AMR_FUSED_LAUNCHES)
        flags = None
        if want_flags:
            loc = torch.empty(n, dtype=torch.uint8, device=dev)
            keep = torch.empty(n, dtype=torch.uint8, device=dev)
            gridpt = torch.empty(n, dtype=torch.uint8, device=dev)
            comp = torch.empty(_lib.MAX_NV, dtype=torch.int32, device=dev)
            _lib.check(lib.sfxm_refine_fused_flags(work.data_ptr(), n, loc.data_ptr(), keep.data_ptr(), gridpt.data_ptr(),
                                                   comp.data_ptr(), st))
            flags = dict(loc=loc, keep=keep, gridpt=gridpt, comp=comp)
        n_new = int(counts[0])          # the one read-back
        return x_buf, U_buf, n_new, flags

    def refine_arrays(self, x, U, want_flags=True):
        """Host arrays in, host arrays out (x[n], U[n, nv]) -> (x_new, U_new, flags as NumPy arrays or None)."""
        import torch
        if not torch.cuda.is_available():
            raise SFXM("splitfxm_b200 needs a CUDA device (there is no CPU fallback)")
        dev = torch.device("cuda", torch.cuda.current_device())
        xd = torch.from_numpy(np.ascontiguousarray(x, dtype=np.float64)).to(dev)
        Ud = torch.from_numpy(np.ascontiguousarray(U, dtype=np.float64)).to(dev)
        x_buf, U_buf, n_new, flags = self._refine_device(xd, Ud, want_flags)
        if flags is not None:
            flags = {k: v.cpu().numpy() for k, v in flags.items()}
        return x_buf[:n_new].cpu().numpy(), U_buf[:n_new].cpu().numpy(), flags

    @staticmethod
    def _dicts_from_flags(flags, names):
        """The reference's `loc`, `c`, `keep` dictionaries (refine.py:103-118) from the device flags; only the indices
        of the marked cells cross to the host.  Keys are in ascending order (the reference's are in the order its loops
        first touch them), `keep` lists the deletions only."""
        import torch
        n = flags["loc"].shape[0]
        loc_idx = torch.nonzero(flags["loc"][:n - 1]).reshape(-1).cpu().tolist()
        del_idx = torch.nonzero(flags["keep"] == 0).reshape(-1).cpu().tolist()
        pt_idx = torch.nonzero(flags["gridpt"]).reshape(-1).cpu().tolist()
        comp = flags["comp"].cpu().tolist()
        loc = {int(i): 1 for i in loc_idx}
        c = {names[i]: 1 for i in range(len(names)) if comp[i]}
        c.update({f"point {int(j)}": 1 for j in pt_idx})
        keep = {int(i): -1 for i in del_idx}
        return loc, c, keep

    def refine(self, d):
        """Refine the Domain's interior in place (refine.py:90-237, 239-282).  Works on the Domain's device copy and
        leaves the new mesh there: nothing but the new size and the indices for the report crosses PCIe."""
        import torch
        if not torch.cuda.is_available():
            raise SFXM("splitfxm_b200 needs a CUDA device (there is no CPU fallback)")
        nbl, nbr, nx = d.nb(btype.LEFT), d.nb(btype.RIGHT), d._nx
        dx, dU = d._device()
        # the kernels write the new interior between the places of the ghost cells: no concatenation afterwards
        x_buf, U_buf, n_new, flags = self._refine_device(dx[nbl:nbl + nx], dU[nbl:nbl + nx], pad=(nbl, nbr))
        names = [d.component_name(i) for i in range(d.num_components())]
        self.show_changes(*self._dicts_from_flags(flags, names))
        n_old = dx.shape[0]
        x_buf[:nbl] = dx[:nbl]
        U_buf[:nbl] = dU[:nbl]
        x_buf[nbl + n_new:nbl + n_new + nbr] = dx[n_old - nbr:]
        U_buf[nbl + n_new:nbl + n_new + nbr] = dU[n_old - nbr:]
        d._adopt_device(x_buf[:nbl + n_new + nbr], U_buf[:nbl + n_new + nbr], nx=n_new, mesh_changed=True)

    def perform_changes(self, d, loc, keep):
        """Apply explicit changes (refine.py:239-282): a midpoint cell after every index in `loc`, every cell whose
        `keep` entry is -1 deleted.  The reference deletes while enumerating the list (refine.py:273-275), so of two
        ADJACENT marked cells it skips the second; reproduced (refine() never marks adjacent cells)."""
        import torch
        if not torch.cuda.is_available():
            raise SFXM("splitfxm_b200 needs a CUDA device (there is no CPU fallback)")
        lib = _lib.load()
        nbl, nbr, nx = d.nb(btype.LEFT), d.nb(btype.RIGHT), d._nx
        marked = np.zeros(nx, dtype=bool)
        for i, v in keep.items():
            if 0 <= i < nx and v == -1:
                marked[i] = True
        # offset of every marked cell inside its run of marked cells: even offsets are deleted, odd ones skipped
        idx = np.arange(nx)
        run_start = np.maximum.accumulate(np.where(marked & ~np.r_[False, marked[:-1]], idx, -1))
        delete = marked & (((idx - run_start) & 1) == 0)
        ins = np.zeros(nx, dtype=np.uint8)
        for i in loc.keys():
            if not 0 <= i < nx - 1:
                raise IndexError("list index out of range")            # cells[i + 1] in the reference
            ins[i] = 1
        dx, dU = d._device()
        dev = dx.device
        xi, Ui = dx[nbl:nbl + nx], dU[nbl:nbl + nx]
        nv = Ui.shape[1]
        work = torch.empty(lib.sfxm_refine_workspace_bytes(nx), dtype=torch.uint8, device=dev)
        counts = torch.zeros(3, dtype=torch.int64, device=dev)
        loc_d = torch.from_numpy(ins).to(dev)
        keep_d = torch.from_numpy((~delete).astype(np.uint8)).to(dev)
        st = _lib.stream_ptr()
        _lib.check(lib.sfxm_refine_set_flags(loc_d.data_ptr(), keep_d.data_ptr(), nx, work.data_ptr(), counts.data_ptr(), st))
        _lib.count(4)
        n_new = int(counts[0])
        x_new = torch.empty(n_new, dtype=torch.float64, device=dev)
        U_new = torch.empty((n_new, nv), dtype=torch.float64, device=dev)
        _lib.check(lib.sfxm_refine_apply(xi.data_ptr(), Ui.data_ptr(), nx, nv, work.data_ptr(), x_new.data_ptr(),
                                         U_new.data_ptr(), st))
        _lib.count()
        n_old = dx.shape[0]
        d._adopt_device(torch.cat([dx[:nbl], x_new, dx[n_old - nbr:]]), torch.cat([dU[:nbl], U_new, dU[n_old - nbr:]]),
                        nx=n_new, mesh_changed=True)

    def show_changes(self, loc, c, keep):
        """The reference's stdout report (refine.py:284-312), same signature: `loc` {index: 1}, `c` {name: 1},
        `keep` {index: +-1}; printed in the dictionaries' own order."""
        print("#" * 78)
        if len(loc) != 0:
            print("Refining grid...")
            print("New points inserted after grid points ")
            for i in loc.keys():
                print(i, end=" ")
            print("    to resolve ", end="")
            for name in c.keys():
                print(name, end=" ")
            print("")
        else:
            print("No new points needed")
        num_deleted = list(keep.values()).count(-1)
        if num_deleted != 0:
            print("Deleted points at ")
            for i in keep.keys():
                if keep[i] == -1:
                    print(i, end=" ")
            print("")
        else:
            print("No new points deleted")
        print("#" * 78)
