"""Slab decomposition of ONE large mesh over the GPUs of a box, with halo exchange (BASELINE config 5).

The reference has no domain decomposition (the whole mesh is one Python list, domain.py:40-41).  Here each rank
owns a contiguous slab of interior cells; the ghost cells a slab needs from its neighbours are ordinary ghost
cells of boundary type SFXM_BC_HALO whose values sit in two small device buffers.  Before every right-hand-side
evaluation each rank sends its first ``nbr`` / last ``nbl`` cells to the left / right neighbour (periodic
variables wrap rank 0 <-> G-1): ``nb * nv`` doubles per side, i.e. latency-bound traffic, sent with
``torch.distributed`` point-to-point operations (NCCL over NVLink on the box, gloo in the CPU tests).  Physical
boundary conditions are applied by the kernels on the ranks that own a domain end, exactly as on one GPU, so
the slab residuals concatenated are bit-identical to the single-mesh residual.
"""
import ctypes

import numpy as np
import torch

from . import _lib
from .bc import _check_dirs, _parse_side
from .engine import Engine, flat_index_table
from .error import SFXM
from .schemes import FVSchemes, _scheme_ids, stencil_sizes


def slab_bounds(nx, world, rank):
    """Balanced contiguous partition: cells [lo, hi) of rank."""
    base, rem = divmod(nx, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def slab_neighbours(rank, world, any_per_left, any_per_right):
    """(left, right) neighbour ranks of a slab, None where there is no neighbour.  Interior faces always have one; the
    domain ends wrap (rank 0 <-> G-1) when some variable is periodic on EITHER side - one symmetric condition, so that
    both end ranks post matching messages when a variable is periodic on one side only (bc.py:50-65 allows
    {'left': 'periodic', 'right': 'outflow'}).  The extra message is harmless: the ghost type of a non-periodic variable
    stays non-HALO and never reads the buffer."""
    wrap = bool(any_per_left or any_per_right)
    left = (rank - 1) if rank > 0 else ((world - 1) if wrap else None)
    right = (rank + 1) if rank < world - 1 else (0 if wrap else None)
    return left, right


class HaloTransport:
    """Moves the packed boundary cells between neighbouring ranks.  send_left goes to the left neighbour and
    lands in ITS recv_right; send_right goes to the right neighbour's recv_left.  `left`/`right` are neighbour
    ranks or None (no neighbour: a non-periodic physical boundary)."""

    def __init__(self, rank, world, left, right, group=None):
        self.rank, self.world, self.left, self.right, self.group = rank, world, left, right, group

    def exchange(self, send_left, send_right, recv_left, recv_right):
        import torch.distributed as dist
        if self.world == 1:
            # periodic wrap onto itself
            if self.left is not None:
                recv_left.copy_(send_right)
            if self.right is not None:
                recv_right.copy_(send_left)
            return
        # Messages between one pair of ranks match in posting order, and with 2 ranks and periodic data the
        # left and the right neighbour are the same rank: post sends as [to-left, to-right] and receives as
        # [from-right, from-left], so a peer's "to-left" message lands in my recv_right.
        ops = []
        if self.left is not None:
            ops.append(dist.P2POp(dist.isend, send_left, self.left, self.group))
        if self.right is not None:
            ops.append(dist.P2POp(dist.isend, send_right, self.right, self.group))
        if self.right is not None:
            ops.append(dist.P2POp(dist.irecv, recv_right, self.right, self.group))
        if self.left is not None:
            ops.append(dist.P2POp(dist.irecv, recv_left, self.left, self.group))
        if ops:
            for req in dist.batch_isend_irecv(ops):
                req.wait()


class LocalTransport:
    """All slabs live in one process (tests, or several slabs per GPU): direct copies between their buffers."""

    def __init__(self):
        self.slabs = []

    def exchange_all(self, packed):
        """packed[k] = (send_left, send_right) of slab k; fills every slab's halo buffers."""
        for k, slab in enumerate(self.slabs):
            if slab.left is not None:
                slab.halo_left.copy_(packed[slab.left][1])      # left neighbour's LAST cells
            if slab.right is not None:
                slab.halo_right.copy_(packed[slab.right][0])    # right neighbour's FIRST cells


class SlabSimulation:
    """One rank's slab of a single mesh.  model: 'ad' | 'euler_prim' | 'euler_cons' (device functors)."""

    def __init__(self, x_global, comps, model, params, bcs, scheme, method, rank, world, xmin=0.0, xmax=1.0,
                 nbl=1, nbr=1, device=None, transport=None, batch=1):
        from .ensemble import resolve_model
        if (method == "FVM") != isinstance(scheme, FVSchemes):
            raise SFXM("Scheme family does not match the numerical method")
        if stencil_sizes.get(scheme) != nbl + nbr + 1:
            raise SFXM("Stencil size does not match boundary conditions")
        x_global = np.ascontiguousarray(x_global, dtype=np.float64)
        nxg = x_global.shape[0]
        self.rank, self.world = rank, world
        self.lo, self.hi = slab_bounds(nxg, world, rank)
        self.nv, self.nbl, self.nbr = len(comps), nbl, nbr
        self.comps = list(comps)
        model_id, cols = resolve_model(model, params)
        table = np.array([float(c) for c in cols])
        self.eng = Engine(model_id, _scheme_ids[scheme], self.nv, nbl, nbr, x_global[self.lo:self.hi], xmin, xmax, table,
                          batch=batch, device=device)
        dev = self.eng.device
        B = batch
        self.halo_left = torch.zeros((nbl, self.nv, B), dtype=torch.float64, device=dev)
        self.halo_right = torch.zeros((nbr, self.nv, B), dtype=torch.float64, device=dev)
        self.send_left = torch.empty((nbr, self.nv, B), dtype=torch.float64, device=dev)    # my FIRST nbr cells
        self.send_right = torch.empty((nbl, self.nv, B), dtype=torch.float64, device=dev)   # my LAST nbl cells
        first, last = rank == 0, rank == world - 1
        # parse the physical BCs once
        parsed = {}
        for name, bc in bcs.items():
            _check_dirs(bc)
            parsed[self.comps.index(name)] = (_parse_side(bc["left"]), _parse_side(bc["right"]))
        any_per_left = any(p[0][0] == _lib.BC_PERIODIC for p in parsed.values())
        any_per_right = any(p[1][0] == _lib.BC_PERIODIC for p in parsed.values())
        self.left, self.right = slab_neighbours(rank, world, any_per_left, any_per_right)
        order = [self.comps.index(n) for n in bcs.keys()]      # application order decides the ghost positions
        nx_loc = self.hi - self.lo

        def fill(P):
            for j in range(self.nv):
                P.bc_ltype[j] = P.bc_rtype[j] = _lib.BC_KEEP
                for g in range(_lib.MAX_NB):
                    P.bc_lval[j][g] = P.bc_rval[j][g] = 0.0
            P.gx_left = P.gx_right = _lib.GX_KEEP
            dx0 = (xmax - xmin) / nxg                             # Domain.from_size initial ghosts, domain.py:87-97
            for g in range(_lib.MAX_NB):
                P.gx_keep_left[g] = xmin - (g + 1) * dx0
                P.gx_keep_right[g] = xmax + (g + 1) * dx0
            for j, ((lt, lv), (rt, rv)) in parsed.items():
                P.bc_ltype[j], P.bc_rtype[j] = lt, rt
                for g in range(_lib.MAX_NB):
                    P.bc_lval[j][g], P.bc_rval[j][g] = lv, rv
            # ---- left face
            if not first:
                for j in range(self.nv):
                    P.bc_ltype[j] = _lib.BC_HALO
                for g in range(nbl):
                    P.gx_keep_left[g] = float(x_global[self.lo - 1 - g])
            else:
                for j, ((lt, _), _r) in parsed.items():
                    if lt == _lib.BC_PERIODIC:
                        P.bc_ltype[j] = _lib.BC_HALO              # values come from rank G-1 (bc.py:57)
                if order:
                    last_lt = parsed[order[-1]][0][0]
                    if last_lt == _lib.BC_PERIODIC:               # bc.py:55-56 on the GLOBAL mesh
                        for g in range(nbl):
                            P.gx_keep_left[g] = xmin - (xmax - float(x_global[nxg - 2 - g]))
                    else:
                        P.gx_left = _lib.GX_MIRROR
            # ---- right face
            if not last:
                for j in range(self.nv):
                    P.bc_rtype[j] = _lib.BC_HALO
                for g in range(nbr):
                    P.gx_keep_right[g] = float(x_global[self.hi + g])
            else:
                for j, (_l, (rt, _)) in parsed.items():
                    if rt == _lib.BC_PERIODIC:
                        P.bc_rtype[j] = _lib.BC_HALO              # values come from rank 0 (bc.py:65)
                if order:
                    last_rt = parsed[order[-1]][1][0]
                    if last_rt == _lib.BC_PERIODIC:               # bc.py:63-64 on the GLOBAL mesh
                        for g in range(nbr):
                            P.gx_keep_right[g] = xmax + (float(x_global[g + 1]) - xmin)
                    else:
                        P.gx_right = _lib.GX_MIRROR
            P.halo_left = self.halo_left.data_ptr()
            P.halo_right = self.halo_right.data_ptr()
            P.peer_left_dst = None
            P.peer_right_dst = None

        self.eng.set_bcs(fill)
        self.transport = transport if transport is not None else HaloTransport(rank, world, self.left, self.right)
        self._pack_idx = {}
        if nx_loc < 2 * (nbl + nbr + 1):
            raise SFXM("slab smaller than twice the stencil")

    # -- halo -------------------------------------------------------------------------------------------
    def _indices(self, split, split_locs):
        key = (bool(split), tuple(split_locs) if split_locs else None)
        hit = self._pack_idx.get(key)
        if hit is None:
            _, segs = self.eng.problem(split, split_locs)
            idx = flat_index_table(self.eng.nx, self.nv, segs)
            first = torch.from_numpy(idx[:self.nbr].reshape(-1)).to(self.eng.device)                  # cells 0..nbr-1
            last = torch.from_numpy(idx[::-1][:self.nbl].reshape(-1).copy()).to(self.eng.device)       # cells nx-1, nx-2, ..
            hit = (first, last)
            self._pack_idx[key] = hit
        return hit

    def pack(self, y, split=False, split_locs=None):
        """Boundary cells of y -> (send_left, send_right), in the receiver's ghost order (g = 0 nearest)."""
        first, last = self._indices(split, split_locs)
        torch.index_select(y, 0, first, out=self.send_left.view(-1, self.eng.batch))
        torch.index_select(y, 0, last, out=self.send_right.view(-1, self.eng.batch))
        return self.send_left, self.send_right

    def exchange(self, y, split=False, split_locs=None):
        sl, sr = self.pack(y, split, split_locs)
        self.transport.exchange(sl, sr, self.halo_left, self.halo_right)

    # -- hot path ---------------------------------------------------------------------------------------
    def residual(self, y, out=None, split=False, split_locs=None, exchange=True):
        if exchange:
            self.exchange(y, split, split_locs)
        return self.eng.residual(y, out, split, split_locs)

    def evolve_euler(self, y, t_diff, max_step, split=False, split_locs=None):
        """The reference's explicit Euler loop (simulation.py:455-463) on the slab; every rank runs the same
        step sequence (dt depends on t only), one halo exchange + one fused launch per step."""
        t_current = 0.0
        other = torch.empty_like(y)
        nsteps = 0
        while t_current < t_diff:
            delta_t = min(t_diff - t_current, max_step)
            self.exchange(y, split, split_locs)
            self.eng.euler_step(y, other, delta_t, split, split_locs)
            y, other = other, y
            t_current += delta_t
            nsteps += 1
        return y, nsteps

    def evolve_rk45(self, y, t_diff, max_step=float("inf"), split=False, split_locs=None, n_global=None):
        """scipy's RK45 (ode.RK45Stepper) on the decomposed mesh: every right-hand side is preceded by a halo
        exchange and the error / initial-step norms are summed over the slabs with one scalar all-reduce each, so all
        ranks take identical steps.  Returns the dense-output end value of this slab."""
        import torch.distributed as dist
        from .ode import RK45Stepper

        def rhs(v, out):
            self.exchange(v, split, split_locs)
            self.eng.residual(v, out, split, split_locs)

        def reduce(local):
            if self.world == 1 or not dist.is_initialized():
                return local
            t = torch.tensor([local], dtype=torch.float64, device=self.eng.device)
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
            return float(t[0])

        if n_global is None:
            n_loc = torch.tensor([self.eng.n], dtype=torch.int64, device=self.eng.device)
            if self.world > 1 and dist.is_initialized():
                dist.all_reduce(n_loc, op=dist.ReduceOp.SUM)
            n_global = int(n_loc[0])
        stepper = RK45Stepper(self.eng, split, split_locs, 1e-3, 1e-6, max_step, rhs=rhs, reduce=reduce, n_global=n_global)
        out = stepper.integrate(y, t_diff)
        return out, stepper.nfev

    # -- Newton / Split-Newton on the decomposed mesh -----------------------------------------------------
    def steady_state(self, y, split=False, split_locs=None, dt0=0.0, dtmax=1.0, armijo=False, bounds=None,
                     maxiter=float("inf"), comm=None):
        """Simulation.steady_state (simulation.py:474-526) for this rank's slab, in place on y [n_local, batch]:
        the ordinary NewtonDriver over spike.SlabNewtonEngine (halo exchange before every residual / Jacobian,
        partitioned block solve with one small all-gather, all-reduced norms), so all ranks iterate in lock-step.
        bounds: [lower, upper] per variable, as in the reference.  Returns the iteration count(s) [batch]."""
        from .newton import NewtonDriver
        from .spike import DistComm, SlabNewtonEngine
        if comm is None:
            comm = DistComm()
        ext = None
        if bounds is not None:
            _, segs = self.eng.problem(split, split_locs)
            nx = self.eng.nx
            ext = [np.concatenate([np.tile(np.asarray(b[o:o + s], dtype=np.float64), nx) for (o, s) in segs])
                   for b in bounds]                                         # extend_bounds, simulation.py:204-264
        drv = NewtonDriver(SlabNewtonEngine(self, comm), split, split_locs, dt0, dtmax, armijo, ext)
        return drv.run(y, maxiter)

    # -- fused halo exchange over peer-mapped memory (NVLink) ---------------------------------------------
    def enable_peer_halos(self, group=None):
        """Put the halo buffers in symmetric memory and map the neighbours' buffers (one process per GPU, NCCL group
        initialised).  After this, evolve_euler_fused() needs no communication call per step: the update kernel
        stores the slab's boundary cells directly into the neighbours' ghost buffers."""
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm_mem
        if self.world < 2:
            raise SFXM("peer halos need at least 2 ranks")
        B, nv = self.eng.batch, self.nv
        self._n_l, self._n_r = self.nbl * nv * B, self.nbr * nv * B
        total = 2 * self._n_l + 2 * self._n_r + 8
        group = group if group is not None else dist.group.WORLD
        buf = symm_mem.empty((total,), dtype=torch.float64, device=self.eng.device)
        buf.zero_()
        hdl = symm_mem.rendezvous(buf, group=group.group_name)
        self._peer_buf, self._peer_hdl = buf, hdl
        self._peer_base = list(hdl.buffer_ptrs)
        self._timeout = torch.zeros(1, dtype=torch.int32, device=self.eng.device)
        torch.cuda.synchronize()
        hdl.barrier()

    def _hl(self, rank, par):      # address of rank's halo_left[par]
        return self._peer_base[rank] + 8 * (par * self._n_l)

    def _hr(self, rank, par):      # address of rank's halo_right[par]
        return self._peer_base[rank] + 8 * (2 * self._n_l + par * self._n_r)

    def _flag(self, rank, which):  # which: 0 = "from left", 1 = "from right"
        return self._peer_base[rank] + 8 * (2 * self._n_l + 2 * self._n_r + which)

    def evolve_euler_fused(self, y, t_diff, max_step, split=False, split_locs=None, use_graph=True, overlap=True):
        """The explicit Euler loop of evolve_euler with the halo exchange fused into the update kernel.

        Step k reads ghost buffers of parity k % 2 and stores its updated boundary cells into the neighbours' buffers
        of parity (k + 1) % 2 (peer memory); a one-thread kernel then publishes the step counter in the neighbours'
        flag words, and the next step starts with a one-thread kernel waiting for both neighbours' counters - so a
        neighbour never overwrites a buffer that is still being read.  The first ghost values come from one ordinary
        exchange.  Pairs of full-size steps are captured once in a CUDA graph (the step number lives in device memory)
        and replayed, so the host issues one graph launch per two steps."""
        lib, eng = self.eng.lib, self.eng
        P, _ = eng.problem(split, split_locs)
        n_l, n_r = self._n_l, self._n_r
        B, nv = eng.batch, self.nv
        hl0 = self._peer_buf[0:n_l].view(self.nbl, nv, B)
        hr0 = self._peer_buf[2 * n_l:2 * n_l + n_r].view(self.nbr, nv, B)
        # flags and the step counter back to zero, then the parity-0 ghosts by the ordinary path
        self._peer_buf[2 * n_l + 2 * n_r:].zero_()
        self._timeout.zero_()
        torch.cuda.synchronize()
        self._peer_hdl.barrier()
        sl, sr = self.pack(y, split, split_locs)
        self.transport.exchange(sl, sr, hl0, hr0)
        torch.cuda.synchronize()
        self._peer_hdl.barrier()
        me, L, R = self.rank, self.left, self.right
        counter = self._peer_base[me] + 8 * (2 * n_l + 2 * n_r + 2)
        from_l = self._flag(me, 0) if L is not None else None
        from_r = self._flag(me, 1) if R is not None else None
        to_l = self._flag(L, 1) if L is not None else None      # my left neighbour sees me on its right
        to_r = self._flag(R, 0) if R is not None else None
        # the reference's step sequence (float-accumulated t, simulation.py:455-463)
        dts, t_current = [], 0.0
        while t_current < t_diff:
            dts.append(min(t_diff - t_current, max_step))
            t_current += dts[-1]
        # two persistent work vectors: a result returned from one call comes back as the next call's state, so the pair
        # (state, other vector) repeats and the captured step pair below can be replayed by later calls
        pool = getattr(self, "_euler_pool", None)
        if pool is None or pool[0].shape != y.shape or pool[0].device != y.device:
            pool = self._euler_pool = (torch.empty_like(y), torch.empty_like(y))
            self._euler_graphs = {}
        other = pool[1] if y.data_ptr() == pool[0].data_ptr() else pool[0]
        bufs = [y, other]
        keep = (P.halo_left, P.halo_right)

        # Overlap (second half of round 2): only the two end warps of the warp-cooperative residual kernel read halo values
        # and feed the neighbours' (sfxm_residual_warps); the interior warps go out on the main stream at once, while a
        # second stream waits for the neighbours' counters, runs the end warps and publishes this rank's counter.  The
        # critical path of a step is then max(interior, wait + ends + signal) instead of their sum.
        nw = 0
        if overlap and eng.batch < 32 and not eng.decoupled and eng.W == 3:
            nw = int(lib.sfxm_residual_warp_count(ctypes.byref(P)))
        overlap = nw > 2
        # high priority: the one-block kernels of the wait / end warps / signal chain must not queue behind the interior
        # kernel's pending CTAs (CUPTI timeline, profiles/scripts/trace_slab.py: at default priority the end warps only
        # started in the interior kernel's last wave, which made the step interior + chain again)
        if overlap and getattr(self, "_euler_aux", None) is None:
            self._euler_aux = torch.cuda.Stream(device=eng.device, priority=-1)
        aux = self._euler_aux if overlap else None

        def set_step(k):
            par, nxt = k % 2, (k + 1) % 2
            P.halo_left, P.halo_right = self._hl(me, par), self._hr(me, par)
            P.peer_left_dst = self._hr(L, nxt) if L is not None else None
            P.peer_right_dst = self._hl(R, nxt) if R is not None else None
            return par, nxt

        def warps(par, nxt, dt, lo, hi):
            _lib.check(lib.sfxm_residual_warps(ctypes.byref(P), _lib.ptr(eng.x_int), _lib.ptr(eng.params), eng.param_stride,
                                               _lib.ptr(bufs[par]), None, _lib.ptr(bufs[nxt]), float(dt), lo, hi,
                                               _lib.stream_ptr()))
            _lib.count()

        def one_step(k, dt):
            par, nxt = set_step(k)
            if overlap:
                main = torch.cuda.current_stream()
                aux.wait_stream(main)                         # fork: the previous step is complete on both streams
                warps(par, nxt, dt, 1, nw - 1)                # interior: no halo value enters
                with torch.cuda.stream(aux):
                    st = _lib.stream_ptr()
                    _lib.check(lib.sfxm_halo_wait_counter(counter, from_l, from_r, self._timeout.data_ptr(), st))
                    warps(par, nxt, dt, -1, nw)               # both end warps in one launch
                    _lib.check(lib.sfxm_halo_signal_counter(counter, to_l, to_r, st))
                    _lib.count(2)
                main.wait_stream(aux)                         # join
                return
            st = _lib.stream_ptr()
            _lib.check(lib.sfxm_halo_wait_counter(counter, from_l, from_r, self._timeout.data_ptr(), st))
            eng.euler_step(bufs[par], bufs[nxt], dt, split, split_locs)
            _lib.check(lib.sfxm_halo_signal_counter(counter, to_l, to_r, st))
            _lib.count(2)

        try:
            k = 0
            n_full = 0
            while n_full < len(dts) and dts[n_full] == max_step:
                n_full += 1
            pairs = n_full // 2 if use_graph else 0
            if pairs >= 2:
                if getattr(self, "_euler_side", None) is None:
                    self._euler_side = torch.cuda.Stream(device=eng.device)
                side = self._euler_side
                side.wait_stream(torch.cuda.current_stream())
                # the captured pair holds the two vectors' addresses, the step size, the layout and this mesh's
                # descriptor by value: replayed by every later call with the same ones (capture + instantiation cost
                # ~2 ms, as much as 50 steps of a 1.25 M-cell slab)
                gkey = (bufs[0].data_ptr(), bufs[1].data_ptr(), float(max_step), bool(split),
                        tuple(split_locs) if split_locs else None, bool(overlap), eng.nx, id(eng.x_int), id(P),
                        self._peer_buf.data_ptr(), tuple(eng._params_host.reshape(-1)))
                graph = self._euler_graphs.get(gkey)
                with torch.cuda.stream(side):
                    if graph is None:
                        graph = torch.cuda.CUDAGraph()
                        with torch.cuda.graph(graph, stream=side):
                            one_step(0, max_step)
                            one_step(1, max_step)
                        self._euler_graphs[gkey] = graph
                    for _ in range(pairs):
                        graph.replay()
                torch.cuda.current_stream().wait_stream(side)
                k = 2 * pairs
            while k < len(dts):
                one_step(k, dts[k])
                k += 1
            torch.cuda.synchronize()
        finally:
            P.halo_left, P.halo_right = keep
            P.peer_left_dst = P.peer_right_dst = None
        if int(self._timeout[0]) != 0:
            raise SFXM("fused halo exchange: a neighbour's step counter did not arrive (timeout)")
        self._peer_hdl.barrier()
        return bufs[len(dts) % 2], len(dts)
