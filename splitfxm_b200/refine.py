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
    def refine_arrays(self, x, U, want_flags=True):
        """Device refinement of one mesh given as host arrays x[n], U[n, nv] -> (x_new, U_new, flags).

        flags = dict(loc, keep, gridpt, comp) as NumPy arrays (for the report), or None."""
        import torch
        if not torch.cuda.is_available():
            raise SFXM("splitfxm_b200 needs a CUDA device (there is no CPU fallback)")
        lib = _lib.load()
        x = np.ascontiguousarray(x, dtype=np.float64)
        U = np.ascontiguousarray(U, dtype=np.float64)
        n, nv = U.shape
        if n > self._npmax:
            raise SFXM("Exceeded maximum number of points")        # refine.py:124-125
        dev = torch.device("cuda", torch.cuda.current_device())
        xd = torch.from_numpy(x).to(dev)
        Ud = torch.from_numpy(U).to(dev)
        work = torch.empty(lib.sfxm_refine_workspace_bytes(n), dtype=torch.uint8, device=dev)
        counts = torch.zeros(3, dtype=torch.int64, device=dev)
        st = _lib.stream_ptr()
        _lib.check(lib.sfxm_refine_mark(xd.data_ptr(), Ud.data_ptr(), n, nv, float(self._ratio), float(self._slope),
                                        float(self._curve), float(self._prune), float(self._min_range),
                                        float(self._min_grid), work.data_ptr(), counts.data_ptr(), st))
        _lib.count(10)
        n_new = int(counts[0])          # the one read-back the new allocation needs
        x_new = torch.empty(n_new, dtype=torch.float64, device=dev)
        U_new = torch.empty((n_new, nv), dtype=torch.float64, device=dev)
        _lib.check(lib.sfxm_refine_apply(xd.data_ptr(), Ud.data_ptr(), n, nv, work.data_ptr(), x_new.data_ptr(),
                                         U_new.data_ptr(), st))
        _lib.count()
        flags = None
        if want_flags:
            loc = torch.empty(n, dtype=torch.uint8, device=dev)
            keep = torch.empty(n, dtype=torch.uint8, device=dev)
            gridpt = torch.empty(n, dtype=torch.uint8, device=dev)
            comp = torch.empty(_lib.MAX_NV, dtype=torch.int32, device=dev)
            _lib.check(lib.sfxm_refine_flags(work.data_ptr(), n, loc.data_ptr(), keep.data_ptr(), gridpt.data_ptr(),
                                             comp.data_ptr(), st))
            flags = dict(loc=loc.cpu().numpy(), keep=keep.cpu().numpy(), gridpt=gridpt.cpu().numpy(),
                         comp=comp.cpu().numpy())
        return x_new.cpu().numpy(), U_new.cpu().numpy(), flags

    def refine(self, d):
        """Refine the Domain's interior in place (refine.py:90-237, 239-282)."""
        nbl, nbr = d.nb(btype.LEFT), d.nb(btype.RIGHT)
        x_new, U_new, flags = self.refine_arrays(d.interior_x(), d.interior_U())
        names = [d.component_name(i) for i in range(d.num_components())]
        self.show_changes(flags, names)
        n_old = d._x.shape[0]
        d._x = np.concatenate([d._x[:nbl], x_new, d._x[n_old - nbr:]])
        d._U = np.concatenate([d._U[:nbl], U_new, d._U[n_old - nbr:]])
        d._nx = x_new.shape[0]
        d._domain = None       # cell views are rebuilt on demand

    def show_changes(self, flags, names):
        """The reference's stdout report (refine.py:284-312); indices are listed in ascending order."""
        print("#" * 78)
        loc = np.nonzero(flags["loc"][:-1])[0]
        if loc.size != 0:
            print("Refining grid...")
            print("New points inserted after grid points ")
            print(" ".join(str(int(i)) for i in loc), end=" ")
            print("    to resolve ", end="")
            why = [names[i] for i in range(len(names)) if flags["comp"][i]]
            why += [f"point {int(j)}" for j in np.nonzero(flags["gridpt"])[0]]
            print(" ".join(why), end=" ")
            print("")
        else:
            print("No new points needed")
        deleted = np.nonzero(flags["keep"] == 0)[0]
        if deleted.size != 0:
            print("Deleted points at ")
            print(" ".join(str(int(i)) for i in deleted), end=" ")
            print("")
        else:
            print("No new points deleted")
        print("#" * 78)
