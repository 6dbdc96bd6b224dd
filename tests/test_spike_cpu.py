"""Host side of the slab-decomposed Newton solve (splitfxm_b200/spike.py) on CPU: the interface system against a dense
solve of the whole block-tridiagonal (cyclic and not) matrix, the in-process thread collectives, and the same
all-gather + interface solve over gloo with world_size 2."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from test_distributed_cpu import _free_port


def _global_system(rng, n, S, periodic, B):
    """Random diagonally dominant block-tridiagonal systems: A[i], Bd[i], C[i] [n, S, S, B], rhs [n, S, B]."""
    A = rng.standard_normal((n, S, S, B))
    C = rng.standard_normal((n, S, S, B))
    Bd = rng.standard_normal((n, S, S, B)) + 6.0 * np.eye(S)[None, :, :, None]
    if not periodic:
        A[0] = 0.0
        C[-1] = 0.0
    return A, Bd, C, rng.standard_normal((n, S, B))


def _dense(A, Bd, C, b):
    n, S = A.shape[0], A.shape[1]
    M = np.zeros((n * S, n * S))
    for i in range(n):
        M[i * S:(i + 1) * S, i * S:(i + 1) * S] = Bd[i, :, :, b]
        jl, jr = (i - 1) % n, (i + 1) % n
        M[i * S:(i + 1) * S, jl * S:(jl + 1) * S] += A[i, :, :, b]
        M[i * S:(i + 1) * S, jr * S:(jr + 1) * S] += C[i, :, :, b]
    return M


def _solve(aug, G, S):
    """Dense NumPy solve of the augmented interface systems [B, N, N + 1] -> z [G, 2, S, B] (on the device the product
    uses csrc dense_solve_kernel; this is the checker for the assembly)."""
    aug = aug.numpy()
    z = np.stack([np.linalg.solve(a[:, :-1], a[:, -1]) for a in aug])            # [B, N]
    return z.reshape(-1, G, 2, S).transpose(1, 2, 3, 0)


def _slab_rows(A, Bd, C, rhs, lo, hi):
    """What sfxm_spike_reduce leaves behind for the slab [lo, hi): x = q - V xl_prev - W xf_next per cell, per member,
    by a dense solve of the slab's own block-tridiagonal matrix with 1 + 2 S right-hand sides."""
    n_loc, S, B = hi - lo, A.shape[1], A.shape[3]
    out = np.zeros((n_loc, S, 1 + 2 * S, B))
    for b in range(B):
        T = np.zeros((n_loc * S, n_loc * S))
        R = np.zeros((n_loc * S, 1 + 2 * S))
        for k in range(n_loc):
            i = lo + k
            T[k * S:(k + 1) * S, k * S:(k + 1) * S] = Bd[i, :, :, b]
            if k > 0:
                T[k * S:(k + 1) * S, (k - 1) * S:k * S] = A[i, :, :, b]
            if k < n_loc - 1:
                T[k * S:(k + 1) * S, (k + 1) * S:(k + 2) * S] = C[i, :, :, b]
            R[k * S:(k + 1) * S, 0] = rhs[i, :, b]
        R[0:S, 1:1 + S] = A[lo, :, :, b]
        R[(n_loc - 1) * S:, 1 + S:] = C[hi - 1, :, :, b]
        out[..., b] = np.linalg.solve(T, R).reshape(n_loc, S, 1 + 2 * S)
    return out


@pytest.mark.parametrize("periodic", [False, True])
@pytest.mark.parametrize("G,S", [(2, 1), (2, 3), (3, 2), (8, 3)])
def test_interface_system_matches_dense_solve(G, S, periodic):
    from splitfxm_b200.distributed import slab_bounds
    from splitfxm_b200.spike import interface_system
    rng = np.random.default_rng(100 * G + S)
    n, B = 5 * G + 3, 2
    A, Bd, C, rhs = _global_system(rng, n, S, periodic, B)
    rows = [_slab_rows(A, Bd, C, rhs, *slab_bounds(n, G, k)) for k in range(G)]
    iface = torch.from_numpy(np.stack([np.stack((r[0], r[-1])) for r in rows]))          # [G, 2, S, 1+2S, B]
    z = _solve(interface_system(iface), G, S)
    for b in range(B):
        ref = np.linalg.solve(_dense(A, Bd, C, b), rhs[:, :, b].reshape(-1)).reshape(n, S)
        got = np.concatenate([r[:, :, 0, b] - r[:, :, 1:1 + S, b] @ z[(k - 1) % G, 1, :, b]
                              - r[:, :, 1 + S:, b] @ z[(k + 1) % G, 0, :, b] for k, r in enumerate(rows)])
        assert np.max(np.abs(got - ref)) < 1e-11 * max(1.0, np.max(np.abs(ref)))
        for k in range(G):
            lo, hi = slab_bounds(n, G, k)
            assert np.allclose(z[k, 0, :, b], ref[lo], rtol=0, atol=1e-11)
            assert np.allclose(z[k, 1, :, b], ref[hi - 1], rtol=0, atol=1e-11)


def test_thread_group_collectives():
    from splitfxm_b200.spike import ThreadGroup, ThreadTransport
    G = 3

    def body(rank, comm):
        t = torch.tensor([1.0 + rank, 10.0 * rank], dtype=torch.float64)
        gathered = comm.all_gather(t)
        total = comm.all_reduce_sum(t.clone())
        low = comm.all_reduce_min(t.clone())
        flag = comm.any(rank == 1)
        got = comm.exchange({(rank + 1) % G: torch.tensor([float(rank)])}, {(rank - 1) % G: (1,)}, t)
        assert float(got[(rank - 1) % G][0]) == float((rank - 1) % G)
        left, right = (rank - 1) % G, (rank + 1) % G
        rl, rr = torch.zeros(1), torch.zeros(1)
        ThreadTransport(comm, left, right).exchange(torch.tensor([float(rank)]), torch.tensor([rank + 0.5]), rl, rr)
        return gathered, total, low, flag, float(rl[0]), float(rr[0])

    out = ThreadGroup(G).run(body)
    for rank, (gathered, total, low, flag, rl, rr) in enumerate(out):
        assert gathered.shape == (G, 2) and gathered[2, 0] == 3.0
        assert total.tolist() == [6.0, 30.0] and low.tolist() == [1.0, 0.0] and flag
        assert rl == (rank - 1) % G + 0.5 and rr == float((rank + 1) % G)


def test_thread_group_reports_failure():
    from splitfxm_b200.spike import ThreadGroup

    def body(rank, comm):
        if rank == 0:
            raise ValueError("boom")
        comm.all_gather(torch.zeros(1))

    with pytest.raises(ValueError, match="boom"):
        ThreadGroup(2, timeout=20.0).run(body)


def _gloo_worker(rank, world, port, periodic, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from splitfxm_b200.distributed import slab_bounds
    from splitfxm_b200.spike import DistComm, interface_system
    rng = np.random.default_rng(5)                       # every rank builds the same global system
    n, S, B = 11, 2, 1
    A, Bd, C, rhs = _global_system(rng, n, S, periodic, B)
    lo, hi = slab_bounds(n, world, rank)
    r = _slab_rows(A, Bd, C, rhs, lo, hi)
    comm = DistComm()
    z = _solve(interface_system(comm.all_gather(torch.from_numpy(np.stack((r[0], r[-1]))))), world, S)
    x = r[:, :, 0, 0] - r[:, :, 1:1 + S, 0] @ z[(rank - 1) % world, 1, :, 0] - r[:, :, 1 + S:, 0] @ z[(rank + 1) % world, 0, :, 0]
    nrm = comm.all_reduce_sum(torch.tensor([float((x ** 2).sum())], dtype=torch.float64))
    low = comm.all_reduce_min(torch.tensor([float(rank)], dtype=torch.float64))
    # the point-to-point range exchange of the slab re-balancing, and the object gather of the refinement report
    other = 1 - rank
    got = comm.exchange({other: torch.full((2 + rank, 3), float(rank))}, {other: (2 + other, 3)}, torch.zeros(1, dtype=torch.float32))
    assert got[other].shape == (2 + other, 3) and bool((got[other] == float(other)).all())
    assert comm.gather_objects(("r", rank)) == [("r", 0), ("r", 1)]
    out[rank] = (lo, hi, x, float(nrm[0]), float(low[0]), comm.any(rank == 1, "cpu"))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("periodic", [False, True])
def test_interface_solve_gloo_world2(periodic):
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_gloo_worker, args=(world, _free_port(), periodic, out), nprocs=world, join=True)
    rng = np.random.default_rng(5)
    A, Bd, C, rhs = _global_system(rng, 11, 2, periodic, 1)
    ref = np.linalg.solve(_dense(A, Bd, C, 0), rhs[:, :, 0].reshape(-1)).reshape(11, 2)
    for rank in range(world):
        lo, hi, x, nrm, low, flag = out[rank]
        assert np.max(np.abs(x - ref[lo:hi])) < 1e-11
        assert abs(nrm - float((ref ** 2).sum())) < 1e-9 and low == 0.0 and flag


# ---- re-balancing of the slabs after a refinement pass (splitfxm_b200/slab_refine.py rebalance_rows) ----
@pytest.mark.parametrize("sizes", [[10, 30, 10, 10], [0, 7, 0, 33], [5, 5, 5, 5], [40, 0, 0, 0], [1, 2, 3], [13]])
def test_rebalance_rows_thread_ranks(sizes):
    from splitfxm_b200.distributed import slab_bounds
    from splitfxm_b200.slab_refine import rebalance_rows
    from splitfxm_b200.spike import ThreadGroup
    G, total = len(sizes), sum(sizes)
    offs = np.concatenate(([0], np.cumsum(sizes)))
    glob = np.stack([np.arange(total, dtype=np.float64), 100.0 + np.arange(total)], axis=1)

    def body(rank, comm):
        mine = torch.from_numpy(glob[offs[rank]:offs[rank + 1]].copy())
        out, tlo, tot = rebalance_rows(comm, mine, sizes)
        return out.numpy(), tlo, tot

    res = ThreadGroup(G).run(body)
    for rank, (out, tlo, tot) in enumerate(res):
        lo, hi = slab_bounds(total, G, rank)
        assert tot == total and tlo == lo
        assert np.array_equal(out, glob[lo:hi])              # the right rows, in order, nothing lost or duplicated


def _rebalance_gloo_worker(rank, world, port, sizes, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from splitfxm_b200.slab_refine import rebalance_rows
    from splitfxm_b200.spike import DistComm
    offs = np.concatenate(([0], np.cumsum(sizes)))
    total = int(offs[-1])
    glob = np.stack([np.arange(total, dtype=np.float64), -np.arange(total, dtype=np.float64)], axis=1)
    mine = torch.from_numpy(glob[offs[rank]:offs[rank + 1]].copy())
    rows, tlo, tot = rebalance_rows(DistComm(), mine, sizes)
    out[rank] = (rows.numpy().copy(), tlo, tot)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("sizes", [[3, 17], [20, 0], [9, 9]])
def test_rebalance_rows_gloo_world2(sizes):
    from splitfxm_b200.distributed import slab_bounds
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_rebalance_gloo_worker, args=(2, _free_port(), sizes, out), nprocs=2, join=True)
    total = sum(sizes)
    for rank in range(2):
        rows, tlo, tot = out[rank]
        lo, hi = slab_bounds(total, 2, rank)
        assert (tlo, tot) == (lo, total)
        assert np.array_equal(rows[:, 0], np.arange(lo, hi, dtype=np.float64))
        assert np.array_equal(rows[:, 1], -np.arange(lo, hi, dtype=np.float64))


def test_refinement_overlap_rows_thread_ranks():
    """The six overlap cells a slab copies from each neighbour before marking (slab_refine._neighbour_rows)."""
    from splitfxm_b200.distributed import slab_bounds
    from splitfxm_b200.slab_refine import OVERLAP, _neighbour_rows
    from splitfxm_b200.spike import ThreadGroup
    G, n = 3, 40
    glob = np.stack([np.arange(n, dtype=np.float64), 2.0 * np.arange(n)], axis=1)

    def body(rank, comm):
        lo, hi = slab_bounds(n, G, rank)
        fl, fr = _neighbour_rows(comm, torch.from_numpy(glob[lo:hi].copy()), lo, hi - lo, n)
        return lo, hi, fl.numpy(), fr.numpy()

    for lo, hi, fl, fr in ThreadGroup(G).run(body):
        assert np.array_equal(fl, glob[max(lo - OVERLAP, 0):lo] if lo > 0 else glob[:0])
        assert np.array_equal(fr, glob[hi:hi + OVERLAP] if hi < n else glob[:0])
