"""Host-side logic of the multi-GPU paths, on CPU with the gloo backend (world_size 2): the halo transport's
send/receive pairing (incl. the periodic wrap where both neighbours are the same rank) and the ensemble
sharding used by bench.py."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _halo_worker(rank, world, port, periodic, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from splitfxm_b200.distributed import HaloTransport
    left = rank - 1 if rank > 0 else (world - 1 if periodic else None)
    right = rank + 1 if rank < world - 1 else (0 if periodic else None)
    tr = HaloTransport(rank, world, left, right)
    # slab k holds cells 10k .. 10k+9 ; first cell / last cell are what travels (nb = 1, nv = 2)
    send_left = torch.tensor([[10.0 * rank, -10.0 * rank]])              # my first cell
    send_right = torch.tensor([[10.0 * rank + 9, -(10.0 * rank + 9)]])   # my last cell
    recv_left = torch.full((1, 2), float("nan"))
    recv_right = torch.full((1, 2), float("nan"))
    tr.exchange(send_left, send_right, recv_left, recv_right)
    out[rank] = (recv_left.numpy().copy(), recv_right.numpy().copy())
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("periodic", [False, True])
def test_halo_transport_world2(periodic):
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    port = _free_port()
    mp.spawn(_halo_worker, args=(world, port, periodic, out), nprocs=world, join=True)
    for rank in range(world):
        rl, rr = out[rank]
        lnb = rank - 1 if rank > 0 else (world - 1 if periodic else None)
        rnb = rank + 1 if rank < world - 1 else (0 if periodic else None)
        if lnb is None:
            assert np.isnan(rl).all()
        else:
            assert rl[0, 0] == 10.0 * lnb + 9          # left halo = left neighbour's LAST cell
        if rnb is None:
            assert np.isnan(rr).all()
        else:
            assert rr[0, 0] == 10.0 * rnb              # right halo = right neighbour's FIRST cell


def _one_sided_worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from splitfxm_b200.distributed import HaloTransport, slab_neighbours
    # a variable periodic on the LEFT only: rank 0 needs rank G-1's last cells; rank G-1 must post the matching send
    left, right = slab_neighbours(rank, world, True, False)
    tr = HaloTransport(rank, world, left, right)
    send_left = torch.tensor([[10.0 * rank]])
    send_right = torch.tensor([[10.0 * rank + 9]])
    recv_left = torch.full((1, 1), float("nan"))
    recv_right = torch.full((1, 1), float("nan"))
    tr.exchange(send_left, send_right, recv_left, recv_right)
    out[rank] = (left, right, float(recv_left[0, 0]), float(recv_right[0, 0]))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_one_sided_periodic_wrap_completes(world):
    """with per-side wrap decisions rank 0 waited for a message rank G-1 never sent."""
    from splitfxm_b200.distributed import slab_neighbours
    assert slab_neighbours(0, 4, False, False) == (None, 1) and slab_neighbours(3, 4, False, False) == (2, None)
    assert slab_neighbours(0, 4, True, False) == (3, 1) and slab_neighbours(3, 4, True, False) == (2, 0)
    assert slab_neighbours(0, 4, False, True) == (3, 1) and slab_neighbours(3, 4, False, True) == (2, 0)
    assert slab_neighbours(0, 1, False, True) == (0, 0)
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_one_sided_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    left, right, rl, rr = out[0]
    assert left == world - 1 and rl == 10.0 * (world - 1) + 9       # rank 0's left ghosts = last cell of rank G-1
    left, right, rl, rr = out[world - 1]
    assert right == 0 and rr == 0.0


def test_slab_bounds_cover_mesh():
    from splitfxm_b200.distributed import slab_bounds
    for nx, world in ((10, 3), (1000, 8), (10_000_000, 8), (17, 1)):
        edges = [slab_bounds(nx, world, r) for r in range(world)]
        assert edges[0][0] == 0 and edges[-1][1] == nx
        assert all(edges[r][1] == edges[r + 1][0] for r in range(world - 1))
        sizes = [b - a for a, b in edges]
        assert max(sizes) - min(sizes) <= 1


def _shard_worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import bench
    B = 8
    c, nu = bench.member_params(rank * B, B, B * world)
    # the timing reduction of bench.py: max over ranks
    t = torch.tensor([1.0 + rank], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    out[rank] = (c.copy(), nu.copy(), float(t[0]))
    dist.destroy_process_group()


def test_ensemble_sharding_world2():
    """Members shard by index with no overlap; every rank's block equals the matching slice of the full sweep."""
    import bench
    world, B = 2, 8
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_shard_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    c_all, nu_all = bench.member_params(0, B * world, B * world)
    for rank in range(world):
        c, nu, tmax = out[rank]
        assert np.array_equal(c, c_all[rank * B:(rank + 1) * B])
        assert np.array_equal(nu, nu_all[rank * B:(rank + 1) * B])
        assert tmax == float(world)
