"""Adaptive mesh refinement of ONE mesh decomposed into slabs.

The reference refines the whole interior in one Python loop (refine.py:90-282).  Here every rank refines its slab
with the same device kernels as the single-GPU Refiner (csrc/sfxm_amr.cu), run on a window = own cells + six overlap
cells copied from each neighbour; what is global in refine.py crosses ranks as three tiny collectives: the min/max
statistics (refine.py:146-153; one 16-double all-reduce), the start of a deletion run that continues from the ranks
before (the "never two adjacent deletions" loop, refine.py:231-233; one int64 all-gather) and the new slab sizes (one
all-gather) - after which contiguous ranges move between ranks so the slabs are balanced again.  The concatenated
result is the reference's refined mesh, bit for bit.
"""
import numpy as np
import torch

from . import _lib
from .distributed import slab_bounds
from .error import SFXM
from .refine import Refiner

OVERLAP = 6      # a rule at cell j reads and flags j-2 .. j+2; the cell before the first own one must be exact too


def _neighbour_rows(comm, rows, lo, n_loc, n_global):
    """Overlap cells from the neighbouring ranks: rows [n_loc, 1 + nv] (x | U) -> (from_left, from_right)."""
    has_l, has_r = lo > 0, lo + n_loc < n_global
    send = {}
    if has_l:
        send[comm.rank - 1] = rows[:OVERLAP].contiguous()
    if has_r:
        send[comm.rank + 1] = rows[n_loc - OVERLAP:].contiguous()
    shape = (OVERLAP, rows.shape[1])
    recv = {}
    if has_l:
        recv[comm.rank - 1] = shape
    if has_r:
        recv[comm.rank + 1] = shape
    got = comm.exchange(send, recv, rows)
    empty = rows[:0]
    return (got[comm.rank - 1] if has_l else empty), (got[comm.rank + 1] if has_r else empty)


def rebalance_rows(comm, mine, sizes):
    """Even out consecutive per-rank row blocks.  Rank r holds `mine` = rows [offs[r], offs[r + 1]) of a global array
    (offs = cumulative `sizes`); afterwards it holds the rows of `slab_bounds(total, world, r)`.  Only the overlaps of the
    old and the new ranges travel, point to point.  Device-agnostic (the CPU tests run it over gloo).
    Returns (rows of the new slab, its first global row, total)."""
    offs = np.concatenate(([0], np.cumsum(sizes)))
    total = int(offs[-1])
    me, world = comm.rank, comm.world
    my_lo, my_hi = int(offs[me]), int(offs[me + 1])
    tlo, thi = slab_bounds(total, world, me)
    send, recv = {}, {}
    for r in range(world):
        if r == me:
            continue
        rlo, rhi = slab_bounds(total, world, r)
        a, b = max(my_lo, rlo), min(my_hi, rhi)              # what I hold of rank r's new slab
        if b > a:
            send[r] = mine[a - my_lo:b - my_lo].contiguous()
        a, b = max(int(offs[r]), tlo), min(int(offs[r + 1]), thi)      # what rank r holds of mine
        if b > a:
            recv[r] = (b - a,) + tuple(mine.shape[1:])
    got = comm.exchange(send, recv, mine)
    parts = []
    for r in range(world):                                   # old ranges are in rank order, so are the pieces
        if r == me:
            a, b = max(my_lo, tlo), min(my_hi, thi)
            if b > a:
                parts.append(mine[a - my_lo:b - my_lo])
        elif r in got:
            parts.append(got[r])
    out = torch.cat(parts, dim=0) if parts else mine[:0]
    return out, tlo, total


class SlabRefiner(Refiner):
    def refine_slab(self, x_loc, U_loc, lo, n_global, comm, rebalance=True, report=True):
        """x_loc [n_loc], U_loc [n_loc, nv]: device tensors with this rank's cells [lo, lo + n_loc) of a mesh of
        n_global cells (ranks hold consecutive ranges in rank order).  Returns (x_new, U_new, lo_new, n_global_new)."""
        if not x_loc.is_cuda:
            raise SFXM("splitfxm_b200 needs a CUDA device (there is no CPU fallback)")
        lib = _lib.load()
        dev = x_loc.device
        n_loc, nv = U_loc.shape
        if n_global > self._npmax:
            raise SFXM("Exceeded maximum number of points")        # refine.py:124-125 (the same on every rank)
        if comm.world > 1 and n_loc < OVERLAP:
            raise SFXM("slab smaller than the refinement overlap (%d cells)" % OVERLAP)
        rows = torch.cat((x_loc.view(-1, 1), U_loc), dim=1).contiguous()
        from_l, from_r = _neighbour_rows(comm, rows, lo, n_loc, n_global)
        ext = torch.cat((from_l, rows, from_r), dim=0)
        ne, own_lo = ext.shape[0], from_l.shape[0]
        own_hi, g0 = own_lo + n_loc, lo - from_l.shape[0]
        xe, Ue = ext[:, 0].contiguous(), ext[:, 1:].contiguous()
        if ne < 3:
            raise SFXM("sfxm_refine_mark: need at least 3 interior cells")
        work = torch.empty(lib.sfxm_refine_workspace_bytes(ne), dtype=torch.uint8, device=dev)
        st = _lib.stream_ptr()
        # 1. statistics: min over ranks of (vmin, -vmax, smin, -smax)
        stats = torch.zeros(4 * _lib.MAX_NV, dtype=torch.float64, device=dev)
        _lib.check(lib.sfxm_refine_window_stats(xe.data_ptr(), Ue.data_ptr(), ne, nv, g0, n_global, own_lo, own_hi,
                                                work.data_ptr(), stats.data_ptr(), st))
        _lib.count(2)
        sign = torch.tensor([1.0, -1.0, 1.0, -1.0], dtype=torch.float64, device=dev).repeat_interleave(_lib.MAX_NV)
        stats = comm.all_reduce_min(stats * sign) * sign
        # 2. flags and the run-start scan; the run that reaches this slab starts at the latest start before it
        last_run = torch.empty(1, dtype=torch.int64, device=dev)
        _lib.check(lib.sfxm_refine_window_mark(xe.data_ptr(), Ue.data_ptr(), ne, nv, g0, n_global, own_lo, own_hi,
                                               float(self._ratio), float(self._slope), float(self._curve),
                                               float(self._prune), float(self._min_range), float(self._min_grid),
                                               stats.data_ptr(), work.data_ptr(), last_run.data_ptr(), st))
        _lib.count(5)
        runs = comm.all_gather(last_run).view(-1).tolist()
        carry = max([-1] + runs[:comm.rank])
        # 3. adjacent-deletion rule, output slots, new sizes
        counts = torch.zeros(3, dtype=torch.int64, device=dev)
        _lib.check(lib.sfxm_refine_window_finalize(ne, g0, n_global, own_lo, own_hi, int(carry), work.data_ptr(),
                                                   counts.data_ptr(), st))
        _lib.count(4)
        sizes = comm.all_gather(counts[:1]).view(-1).tolist()
        n_new = int(sizes[comm.rank])
        x_new = torch.empty(n_new, dtype=torch.float64, device=dev)
        U_new = torch.empty((n_new, nv), dtype=torch.float64, device=dev)
        _lib.check(lib.sfxm_refine_window_apply(xe.data_ptr(), Ue.data_ptr(), ne, nv, g0, n_global, own_lo, own_hi,
                                                work.data_ptr(), x_new.data_ptr(), U_new.data_ptr(), st))
        _lib.count()
        if report:
            self._report(comm, work, ne, own_lo, own_hi, lo, n_global, dev)
        offs = np.concatenate(([0], np.cumsum(sizes)))
        total = int(offs[-1])
        if not rebalance or comm.world == 1:
            return x_new, U_new, int(offs[comm.rank]), total
        # 4. rebalance: contiguous ranges of the new global order move to the ranks that own them now
        mine = torch.cat((x_new.view(-1, 1), U_new), dim=1).contiguous()
        out, tlo, total = rebalance_rows(comm, mine, sizes)
        return out[:, 0].contiguous(), out[:, 1:].contiguous(), tlo, total

    def _report(self, comm, work, ne, own_lo, own_hi, lo, n_global, dev):
        """The reference's stdout report (refine.py:284-312) with global indices, printed by rank 0."""
        lib = _lib.load()
        n = own_hi - own_lo
        loc = torch.empty(n, dtype=torch.uint8, device=dev)
        keep = torch.empty(n, dtype=torch.uint8, device=dev)
        gridpt = torch.empty(n, dtype=torch.uint8, device=dev)
        comp = torch.empty(_lib.MAX_NV, dtype=torch.int32, device=dev)
        _lib.check(lib.sfxm_refine_window_flags(work.data_ptr(), ne, own_lo, own_hi, loc.data_ptr(), keep.data_ptr(),
                                                gridpt.data_ptr(), comp.data_ptr(), _lib.stream_ptr()))
        pieces = comm.gather_objects((lo, loc.cpu().numpy(), keep.cpu().numpy(), gridpt.cpu().numpy(), comp.cpu().numpy()))
        if comm.rank != 0:
            return
        pieces.sort(key=lambda p: p[0])
        flags = dict(loc=np.concatenate([p[1] for p in pieces]), keep=np.concatenate([p[2] for p in pieces]),
                     gridpt=np.concatenate([p[3] for p in pieces]), comp=np.max([p[4] for p in pieces], axis=0))
        names = getattr(self, "names", None) or [f"component {i}" for i in range(_lib.MAX_NV)]
        self.show_changes(*self._dicts_from_flags({k: torch.from_numpy(np.ascontiguousarray(v)) for k, v in flags.items()}, names))
