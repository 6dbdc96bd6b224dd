"""Newton / Split-Newton on ONE mesh decomposed into slabs.

The reference solves the whole mesh in one process (simulation.py:474-526 calling the third-party splitnewton).
Here every rank assembles the Jacobian of its slab (the couplings to the neighbouring slabs sit in the wrap slots,
csrc/sfxm_kernels.cuh jacobian_cell), eliminates its interior with block PCR carrying 1 + 2 s right-hand sides
(csrc/sfxm_spike.cu), all-gathers its two interface rows - 2 s (1 + 2 s) doubles per member, the only data-path
collective of the solve - and every rank solves the same 2 G s interface system before folding the result back into
its slab.  The Newton control (oracle/newton.py's specification) is the ordinary ``NewtonDriver``: its norms become
all-reduced sums, so every rank takes identical decisions.

``Comm`` is the small collective interface the driver needs.  ``DistComm`` maps it onto ``torch.distributed`` (NCCL
on the box, gloo in the CPU tests); ``ThreadGroup`` runs G slabs as G threads of one process sharing one GPU, which
is how the single-GPU test-suite exercises the whole path.
"""
import ctypes
import threading

import torch

from . import _lib
from .error import SFXM


# -- collectives ---------------------------------------------------------------------------------------------
class DistComm:
    def __init__(self, group=None):
        import torch.distributed as dist
        self.dist, self.group = dist, group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)

    def all_gather(self, t):
        parts = [torch.empty_like(t) for _ in range(self.world)]
        self.dist.all_gather(parts, t.contiguous(), group=self.group)
        return torch.stack(parts)

    def all_reduce_sum(self, t):
        self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM, group=self.group)
        return t

    def all_reduce_min(self, t):
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MIN, group=self.group)
        return t

    def gather_objects(self, obj):
        out = [None] * self.world
        self.dist.all_gather_object(out, obj, group=self.group)
        return out

    def exchange(self, send, recv_shapes, like):
        """Point-to-point: send {dst: tensor}, recv_shapes {src: shape} -> {src: tensor} (dtype/device of `like`)."""
        got = {src: torch.empty(shape, dtype=like.dtype, device=like.device) for src, shape in recv_shapes.items()}
        ops = [self.dist.P2POp(self.dist.isend, send[dst], dst, self.group) for dst in sorted(send)]
        ops += [self.dist.P2POp(self.dist.irecv, got[src], src, self.group) for src in sorted(got)]
        if ops:
            for req in self.dist.batch_isend_irecv(ops):
                req.wait()
        return got

    def any(self, flag, device):
        t = torch.tensor([1 if flag else 0], dtype=torch.int32, device=device)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX, group=self.group)
        return bool(int(t[0]))


class ThreadGroup:
    """G ranks as G threads of one process (one GPU): collectives are list hand-offs between two barriers."""

    def __init__(self, world, timeout=120.0):
        self.world = world
        self.barrier = threading.Barrier(world, timeout=timeout)
        self.slots = [None] * world

    def comm(self, rank):
        return ThreadComm(self, rank)

    def run(self, fn):
        """fn(rank, comm) on every rank; returns the list of results, re-raises the first failure."""
        out, err = [None] * self.world, [None] * self.world

        def body(r):
            try:
                out[r] = fn(r, self.comm(r))
            except BaseException as e:       # noqa: BLE001 - reported to the caller below
                err[r] = e
                self.barrier.abort()
        threads = [threading.Thread(target=body, args=(r,)) for r in range(self.world)]
        for t in threads:
            t.start()
        for t in threads:
            t.join()
        real = [e for e in err if e is not None and not isinstance(e, threading.BrokenBarrierError)]
        if real:
            raise real[0]
        for e in err:
            if e is not None:
                raise e
        return out


class ThreadComm:
    def __init__(self, group, rank):
        self.g, self.rank, self.world = group, rank, group.world

    def gather_objects(self, obj):
        self.g.slots[self.rank] = obj
        self.g.barrier.wait()
        out = list(self.g.slots)
        self.g.barrier.wait()
        return out

    def all_gather(self, t):
        return torch.stack(self.gather_objects(t.clone()))

    def exchange(self, send, recv_shapes, like):
        posted = self.gather_objects({dst: t.clone() for dst, t in send.items()})
        return {src: posted[src][self.rank] for src in recv_shapes}

    def all_reduce_sum(self, t):
        parts = self.gather_objects(t.clone())
        acc = parts[0].clone()
        for p in parts[1:]:
            acc += p
        t.copy_(acc)
        return t

    def all_reduce_min(self, t):
        t.copy_(torch.stack(self.gather_objects(t.clone())).min(dim=0).values)
        return t

    def any(self, flag, device=None):
        return any(self.gather_objects(bool(flag)))


class ThreadTransport:
    """HaloTransport twin for ThreadGroup ranks (same argument meaning as HaloTransport.exchange)."""

    def __init__(self, comm, left, right):
        self.comm, self.left, self.right = comm, left, right

    def exchange(self, send_left, send_right, recv_left, recv_right):
        packed = self.comm.gather_objects((send_left.clone(), send_right.clone()))
        if self.left is not None:
            recv_left.copy_(packed[self.left][1])       # left neighbour's LAST cells
        if self.right is not None:
            recv_right.copy_(packed[self.right][0])     # right neighbour's FIRST cells


# -- the interface system --------------------------------------------------------------------------------------
_IFACE_INDEX = {}


def _iface_index(G, S, device):
    """(rows, cols of V entries, cols of W entries) of the interface matrix, entry order (k, e, r, c)."""
    key = (G, S, str(device))
    hit = _IFACE_INDEX.get(key)
    if hit is None:
        k = torch.arange(G).view(G, 1, 1, 1)
        e = torch.arange(2).view(1, 2, 1, 1)
        r = torch.arange(S).view(1, 1, S, 1)
        c = torch.arange(S).view(1, 1, 1, S)
        rows = ((k * 2 + e) * S + r).expand(G, 2, S, S)
        cols_v = ((((k - 1) % G) * 2 + 1) * S + c).expand(G, 2, S, S)     # xl of the left neighbour
        cols_w = ((((k + 1) % G) * 2) * S + c).expand(G, 2, S, S)         # xf of the right neighbour
        hit = tuple(t.reshape(-1).to(device) for t in (rows, cols_v, cols_w))
        _IFACE_INDEX[key] = hit
    return hit


def interface_system(iface_all):
    """iface_all [G, 2, s, 1 + 2 s, B]: per slab the rows (q | V | W) of its first (e = 0) and last (e = 1) cell,
    x = q - V xl_prev - W xf_next.  Returns the augmented matrices [B, N, N + 1], N = 2 G s, of
        xf_k + V_k[0] xl_(k-1) + W_k[0] xf_(k+1) = q_k[0],   xl_k + V_k[1] xl_(k-1) + W_k[1] xf_(k+1) = q_k[1]
    with unknown order (k, e, component).  Slab neighbours wrap around (mod G); a non-periodic domain end simply has
    V = 0 / W = 0."""
    G, two, S, NR, B = iface_all.shape
    if two != 2 or NR != 1 + 2 * S:
        raise SFXM("interface_system: bad interface block shape")
    N = 2 * G * S
    rows, cols_v, cols_w = _iface_index(G, S, iface_all.device)
    aug = torch.zeros((B, N, N + 1), dtype=iface_all.dtype, device=iface_all.device)
    diag = torch.arange(N, device=iface_all.device)
    aug[:, diag, diag] = 1.0
    V = iface_all[:, :, :, 1:1 + S, :].permute(4, 0, 1, 2, 3).reshape(B, -1)
    W = iface_all[:, :, :, 1 + S:, :].permute(4, 0, 1, 2, 3).reshape(B, -1)
    b = torch.arange(B, device=iface_all.device).view(B, 1)
    aug.index_put_((b, rows.view(1, -1), cols_v.view(1, -1)), V, accumulate=True)
    aug.index_put_((b, rows.view(1, -1), cols_w.view(1, -1)), W, accumulate=True)
    aug[:, :, N] = iface_all[:, :, :, 0, :].permute(3, 0, 1, 2).reshape(B, N)
    return aug


def solve_interfaces(iface_all, status=None):
    """-> z [G, 2, s, B], the first / last block unknown of every slab (device tensors: csrc dense_solve_kernel)."""
    G, _, S, _, B = iface_all.shape
    if not iface_all.is_cuda:
        raise SFXM("solve_interfaces needs device tensors (there is no CPU fallback)")
    aug = interface_system(iface_all).contiguous()
    N = 2 * G * S
    z = torch.empty((B, N), dtype=torch.float64, device=iface_all.device)
    _lib.check(_lib.load().sfxm_dense_solve(_lib.ptr(aug), N, B, _lib.ptr(z), _lib.ptr(status), _lib.stream_ptr()))
    _lib.count()
    return z.reshape(B, G, 2, S).permute(1, 2, 3, 0).contiguous()


# -- the engine the Newton driver sees -----------------------------------------------------------------------
class SlabNewtonEngine:
    """Engine facade over one slab: residual / Jacobian are preceded by a halo exchange, the linear solve is the
    partitioned one, sums of squares are summed over the slabs."""

    def __init__(self, slab, comm):
        if slab.eng.W != 3:
            raise SFXM("slab Newton needs a 3-cell stencil (nb 1/1)")
        self.slab, self.eng, self.comm = slab, slab.eng, comm
        self.lib = slab.eng.lib

    # plain attributes of Engine that NewtonDriver reads
    nx = property(lambda self: self.eng.nx)
    n = property(lambda self: self.eng.n)
    batch = property(lambda self: self.eng.batch)
    device = property(lambda self: self.eng.device)

    def problem(self, split=False, split_locs=None):
        return self.eng.problem(split, split_locs)

    def residual(self, y, out=None, split=False, split_locs=None):
        return self.slab.residual(y, out, split, split_locs, exchange=True)

    def jacobian(self, y, split=False, split_locs=None, eps=1e-8, J=None, r0=None, seg_lo=None, seg_hi=None):
        self.slab.exchange(y, split, split_locs)
        return self.eng.jacobian(y, split, split_locs, eps, J=J, r0=r0, seg_lo=seg_lo, seg_hi=seg_hi)

    def solve(self, J, f, s, split=False, split_locs=None, seg_lo=0, seg_hi=None, shift=None, active=None, status=None,
              method=None):
        eng, lib = self.eng, self.lib
        P, segs = eng.problem(split, split_locs)
        if seg_hi is None:
            seg_hi = len(segs)
        B = eng.batch
        work = eng._scratch("spike", lib.sfxm_spike_workspace_bytes(ctypes.byref(P)) // 8)
        for k in range(seg_lo, seg_hi):
            S = segs[k][1]
            iface = torch.empty((2, S, 1 + 2 * S, B), dtype=torch.float64, device=eng.device)
            _lib.check(lib.sfxm_spike_reduce(ctypes.byref(P), _lib.ptr(J), _lib.ptr(f), _lib.ptr(shift), int(k),
                                             _lib.ptr(work), _lib.ptr(iface), _lib.ptr(status), _lib.stream_ptr()))
            _lib.count(2 + max(eng.nx - 1, 1).bit_length())
            z = solve_interfaces(self.comm.all_gather(iface), status)
            G, me = self.comm.world, self.comm.rank
            nbr = torch.stack((z[(me - 1) % G, 1], z[(me + 1) % G, 0])).contiguous()      # [2, S, B]
            _lib.check(lib.sfxm_spike_fold(ctypes.byref(P), int(k), _lib.ptr(work), _lib.ptr(nbr), _lib.ptr(active),
                                           _lib.ptr(s), _lib.stream_ptr()))
            _lib.count()
        return s

    def sumsq(self, v, lo=0, hi=None, out=None):
        return self.comm.all_reduce_sum(self.eng.sumsq(v, lo, hi, out))

    def update_norms(self, y, s, active=None, lo=0, hi=None, update=True, alpha=None):
        ss, yy = self.eng.update_norms(y, s, active, lo, hi, update, alpha)
        both = self.comm.all_reduce_sum(torch.cat((ss, yy)))
        return both[:self.eng.batch], both[self.eng.batch:]

    def axpy(self, y, s, alpha=None, active=None, lo=0, hi=None):
        return self.eng.axpy(y, s, alpha, active, lo, hi)

    # reductions of host-side decisions (bounds)
    def reduce_min(self, t):
        return self.comm.all_reduce_min(t)

    def reduce_any(self, flag):
        return self.comm.any(flag, self.eng.device)
