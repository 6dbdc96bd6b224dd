"""Host side of the Newton / Split-Newton driver (splitfxm_b200/newton.py) on CPU: the driver's control logic - member
masks, segment recursion, pseudo-transient step growth, Armijo back-tracking, bound damping, the criterion - runs on CPU
tensors over a stand-in engine whose residual / Jacobian / linear solve come from the oracle, and must follow the
specification in oracle/newton.py step for step.  (The device engine is tested against the same specification in
tests/test_gpu_newton.py; this covers the host logic where there is no GPU, including the slab variant whose norms are
summed over thread ranks and over gloo.)"""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from test_distributed_cpu import _free_port

COMPS = ["u", "v", "w"]
BCS = {"u": {"left": {"dirichlet": 1.0}, "right": {"dirichlet": 0.0}},
       "v": {"left": {"dirichlet": 3}, "right": {"dirichlet": 4}},
       "w": {"left": {"dirichlet": 2}, "right": {"neumann": 0.0}}}
NX = 24


def _problem():
    from oracle import core
    x = np.linspace(0.0, 1.0, NX)
    P = core.Problem(core.AdvectionDiffusion(0.2, 0.05), "FVM", "LAX_FRIEDRICHS", x, 1, 1, BCS, COMPS)
    U = np.stack([core.initial_condition(x, ic) for ic in ("gaussian", "rarefaction", "tophat")], axis=1)
    return P, U


class OracleEngine:
    """The Engine interface NewtonDriver uses, batch 1, on CPU tensors; rows [row_lo, row_hi) of every flat vector are
    this rank's (the whole vector when there is one rank)."""

    def __init__(self, P, split, locs, comm=None, nx=NX, nv=3):
        from splitfxm_b200.engine import split_segments
        self.P, self.split, self.locs = P, split, locs
        self.nx, self.nv, self.batch = nx, nv, 1
        self.n = nx * nv
        self.device = torch.device("cpu")
        self.segs = split_segments(nv, split, locs)
        self.comm = comm
        if comm is not None:                       # norms are summed over the ranks; every rank holds the whole state
            self.reduce_min = lambda t: comm.all_reduce_min(t)
            self.reduce_any = lambda flag: comm.any(flag, "cpu")

    def problem(self, split=False, split_locs=None):
        return None, self.segs

    def residual(self, y, out=None, split=False, split_locs=None):
        out = torch.empty_like(y) if out is None else out
        out[:, 0] = torch.from_numpy(self.P.residual(y[:, 0].numpy(), self.split, self.locs))
        return out

    def jacobian(self, y, split=False, split_locs=None, eps=1e-8, J=None, r0=None, seg_lo=None, seg_hi=None):
        x = y[:, 0].numpy()
        if r0 is not None:
            r0[:, 0] = torch.from_numpy(self.P.residual(x, self.split, self.locs))
        A = self.P.jacobian_matrix(x, self.split, self.locs)
        return np.asarray(A.todense()) if hasattr(A, "todense") else np.asarray(A)

    def solve(self, J, f, s, split=False, split_locs=None, seg_lo=0, seg_hi=None, shift=None, active=None, status=None,
              method=None):
        for k in range(seg_lo, len(self.segs) if seg_hi is None else seg_hi):
            o, sz = self.segs[k]
            lo, hi = o * self.nx, (o + sz) * self.nx
            A = J[lo:hi, lo:hi] + (float(shift[0]) if shift is not None else 0.0) * np.eye(hi - lo)
            if active is None or int(active[0]):
                s[lo:hi, 0] = torch.from_numpy(np.linalg.solve(A, -f[lo:hi, 0].numpy()))
        return s

    def _share(self, lo, hi):
        """This rank's part of [lo, hi): a contiguous piece, so that the all-reduced sums add up to the whole norm."""
        if self.comm is None:
            return lo, hi
        from splitfxm_b200.distributed import slab_bounds
        a, b = slab_bounds(hi - lo, self.comm.world, self.comm.rank)
        return lo + a, lo + b

    def sumsq(self, v, lo=0, hi=None, out=None):
        hi = self.n if hi is None else hi
        a, b = self._share(lo, hi)
        t = (v[a:b] ** 2).sum(dim=0)
        return self.comm.all_reduce_sum(t) if self.comm is not None else t

    def update_norms(self, y, s, active=None, lo=0, hi=None, update=True, alpha=None):
        hi = self.n if hi is None else hi
        if update and (active is None or int(active[0])):
            y[lo:hi] += (alpha if alpha is not None else 1.0) * s[lo:hi]
        return self.sumsq(s, lo, hi), self.sumsq(y, lo, hi)

    def axpy(self, y, s, alpha=None, active=None, lo=0, hi=None):
        hi = self.n if hi is None else hi
        if active is None or int(active[0]):
            y[lo:hi] += (alpha if alpha is not None else 1.0) * s[lo:hi]
        return y


CASES = [
    (False, None, {}),
    (True, [1, 2], {}),
    (True, [1], {"dt0": 0.05, "dtmax": 10.0}),
    (False, None, {"armijo": True}),
    (True, [2], {"bounds": [[-1.0, -1.0, -1.0], [5.0, 5.0, 5.0]]}),
    (True, [1, 2], {"dt0": 0.02, "dtmax": 1.0, "armijo": True}),
]


def _reference_run(P, U, split, locs, kw):
    from oracle import core, newton as onewton
    x0 = core.to_flat(U, split, locs)
    f = lambda u: P.residual(u, split, locs)
    jac = lambda u: P.jacobian_matrix(u, split, locs)
    okw = dict(kw)
    if "bounds" in okw:
        okw["bounds"] = _extend_bounds(okw["bounds"], split, locs)
    if split:
        xf, _, it = onewton.split_newton(f, jac, x0, [NX * i for i in locs], sparse=True, **okw)
    else:
        xf, _, it = onewton.newton(f, jac, x0, sparse=True, **okw)
    return x0, xf, it


def _extend_bounds(bounds, split, locs):
    from splitfxm_b200.engine import split_segments
    segs = split_segments(3, split, locs)
    return [np.concatenate([np.tile(np.asarray(b[o:o + s], dtype=np.float64), NX) for (o, s) in segs]).tolist()
            for b in bounds]


def _driver_run(eng, x0, split, locs, kw, ext_bounds=None):
    from splitfxm_b200.newton import NewtonDriver
    dkw = dict(kw)
    bounds = dkw.pop("bounds", None)
    ext = ext_bounds if ext_bounds is not None else (_extend_bounds(bounds, split, locs) if bounds is not None else None)
    y = torch.from_numpy(x0.copy()).reshape(-1, 1)
    drv = NewtonDriver(eng, split, locs, dkw.get("dt0", 0.0), dkw.get("dtmax", 1.0), dkw.get("armijo", False), ext)
    assert not drv.fused
    it = drv.run(y)
    return int(it[0]), y[:, 0].numpy()


@pytest.mark.parametrize("split,locs,kw", CASES)
def test_driver_follows_the_specification(split, locs, kw):
    P, U = _problem()
    x0, xf, it_ref = _reference_run(P, U, split, locs, kw)
    it, x = _driver_run(OracleEngine(P, split, locs), x0, split, locs, kw)
    assert it == it_ref
    assert np.max(np.abs(x - xf)) <= 1e-10 * max(1.0, np.max(np.abs(xf)))


def _nonlinear_names():
    from nonlinear_cases import CASES as NL
    return sorted(NL)


@pytest.mark.parametrize("name", _nonlinear_names())
def test_driver_follows_the_specification_on_nonlinear_problems(name):
    """Burgers (viscous shock, custom model) and the Euler duct: >= 4 iterations with active Armijo back-tracking,
    bound clipping and pseudo-transient growth (tests/test_newton_nonlinear_oracle.py asserts those branches are taken);
    the device driver's control logic must reproduce the specification's iteration count and iterate."""
    from nonlinear_cases import extend_bounds, oracle_case, oracle_solve
    P, U0, c = oracle_case(name)
    nx, nv = U0.shape
    x0, xf, it_ref, _ = oracle_solve(name)
    ext = extend_bounds(c["kw"]["bounds"], nx, nv, c["split"], c["locs"]) if "bounds" in c["kw"] else None
    it, x = _driver_run(OracleEngine(P, c["split"], c["locs"], nx=nx, nv=nv), x0, c["split"], c["locs"], c["kw"], ext)
    assert it == it_ref
    assert np.max(np.abs(x - xf)) <= 1e-10 * max(1.0, np.max(np.abs(xf)))


def test_driver_rejects_an_initial_guess_outside_the_bounds():
    P, U = _problem()
    x0, _, _ = _reference_run(P, U, False, None, {})
    with pytest.raises(ValueError, match="outside the bounds"):
        _driver_run(OracleEngine(P, False, None), x0, False, None, {"bounds": [[0.5, 0.5, 0.5], [5.0, 5.0, 5.0]]})


@pytest.mark.parametrize("split,locs,kw", [CASES[1], CASES[2], CASES[4]])
def test_driver_with_norms_summed_over_thread_ranks(split, locs, kw):
    """The slab variant: every reduction of a host-side decision goes through the communicator, so all ranks take the
    same iterations (here three thread ranks share the norm sums of one replicated state)."""
    from splitfxm_b200.spike import ThreadGroup
    P, U = _problem()
    x0, xf, it_ref = _reference_run(P, U, split, locs, kw)

    def body(rank, comm):
        return _driver_run(OracleEngine(P, split, locs, comm), x0, split, locs, kw)

    for it, x in ThreadGroup(3).run(body):
        assert it == it_ref
        assert np.max(np.abs(x - xf)) <= 1e-10 * max(1.0, np.max(np.abs(xf)))


def _gloo_worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from splitfxm_b200.spike import DistComm
    P, U = _problem()
    split, locs, kw = CASES[2]
    x0, xf, it_ref = _reference_run(P, U, split, locs, kw)
    it, x = _driver_run(OracleEngine(P, split, locs, DistComm()), x0, split, locs, kw)
    out[rank] = (it, it_ref, float(np.max(np.abs(x - xf))))
    dist.barrier()
    dist.destroy_process_group()


def test_driver_with_norms_summed_over_gloo_world2():
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_gloo_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    for rank in range(2):
        it, it_ref, err = out[rank]
        assert it == it_ref and err < 1e-9
