#!/usr/bin/env python
"""bench_slab.py - BASELINE config 5: one conservative Sod mesh, slab-decomposed over N GPUs with halo exchange.

Launch:  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
             bench_slab.py --nx 10000000 --steps 200
Every rank owns nx/N cells of the Euler1DConservative FVM Lax-Friedrichs problem (neumann-0 ends,
tutorials/conservative_sod_shock/conservative_sod_shock.py:17-34) and runs explicit Euler steps with
max_step = 0.4 dx / 2.2: per step one NCCL halo exchange (nb*nv doubles per side) and one fused
residual+update launch.  Strong scaling: the mesh is fixed, N varies.  `--check` first verifies, on a small
mesh, that the gathered slabs equal a single-GPU run bit for bit.
Prints one JSON line (rank 0): cell-updates/s, ms per step, achieved HBM GB/s (56 B per cell-step) per GPU.
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

COMPS = ["rho", "rhou", "E"]
BCS = {n: {"left": {"neumann": 0.0}, "right": {"neumann": 0.0}} for n in COMPS}


def sod_state(x):
    left = x < 0.5
    return np.stack([np.where(left, 1.0, 0.125), np.zeros_like(x), np.where(left, 2.5, 0.25)], axis=1)


AD_COMPS = ["u", "v", "w"]
AD_BCS = {"u": {"left": {"dirichlet": 1.0}, "right": {"dirichlet": 0.0}},
          "v": {"left": {"dirichlet": 3}, "right": {"dirichlet": 4}},
          "w": {"left": {"dirichlet": 2}, "right": {"neumann": 0.0}}}
AD_ICS = {"u": "gaussian", "v": "rarefaction", "w": "tophat"}


def newton_on_slabs(args, rank, world, dev):
    """Split-Newton on a slab-decomposed advection-diffusion mesh: parity against the single-GPU
    Simulation.steady_state on a small mesh, then the time per Newton iteration at --newton-nx.  Nothing raises on
    one rank only (the others would wait in the next collective): the outcome is reported."""
    import torch
    import torch.distributed as dist
    import splitfxm_b200 as sf
    from splitfxm_b200.engine import flat_index_table, split_segments
    from splitfxm_b200.distributed import SlabSimulation
    from splitfxm_b200.models.advection_diffusion import AdvectionDiffusion
    from splitfxm_b200.spike import DistComm, ThreadGroup
    comm = DistComm() if world > 1 else ThreadGroup(1).comm(0)

    def to_flat(block, split, locs):          # [nx, nv] -> the flat (split-major) vector, simulation.py:137-143
        out = np.empty(block.size)
        out[flat_index_table(block.shape[0], block.shape[1], split_segments(block.shape[1], split, locs))] = block
        return out

    def to_block(flat, nx, nv, split, locs):
        return np.asarray(flat)[flat_index_table(nx, nv, split_segments(nv, split, locs))]
    scheme = sf.FVSchemes.LAX_FRIEDRICHS

    def setup(nx):
        d = sf.Domain.from_size(nx, 1, 1, AD_COMPS)
        sim = sf.Simulation(d, AdvectionDiffusion(0.2, 0.05, "FVM"), AD_ICS, AD_BCS, scheme)
        xg = np.array(d.positions(interior=True))
        slab = SlabSimulation(xg, AD_COMPS, "ad", {"c": 0.2, "nu": 0.05}, AD_BCS, scheme, "FVM", rank, world, device=dev)
        return d, sim, slab

    out = {}
    # -- parity on a small mesh
    nx, split, locs = 20000, True, [1, 2]        # slabs of >= 2048 cells up to 8 ranks: the nested-dissection path
    d, sim, slab = setup(nx)
    U0 = d.values(interior=True).copy()
    y = slab.eng.to_device(to_flat(U0[slab.lo:slab.hi], split, locs))
    its = slab.steady_state(y, split, locs, comm=comm)
    mine = (slab.lo, slab.hi, int(its[0]), to_block(y.cpu().numpy()[:, 0], slab.hi - slab.lo, 3, split, locs))
    parts = [mine]
    if world > 1:
        parts = [None] * world
        dist.all_gather_object(parts, mine)
    if rank == 0:
        it_ref = sim.steady_state(split=split, split_locs=locs)
        ref = d.values(interior=True)
        U = np.concatenate([p[3] for p in parts])
        err = float(np.max(np.abs(U - ref)) / np.max(np.abs(ref)))
        out["parity"] = {"nx": nx, "iterations": [p[2] for p in parts], "iterations_single_gpu": it_ref,
                         "max_rel_diff_vs_single_gpu": err,
                         "ok": bool(err < 1e-10 and all(p[2] == it_ref for p in parts))}
    # -- timing
    nx = args.newton_nx
    d, sim, slab = setup(nx)
    y0 = slab.eng.to_device(to_flat(d.values(interior=True)[slab.lo:slab.hi], False, None))
    for timed in (False, True):
        y = y0.clone()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record()
        its = slab.steady_state(y, False, None, comm=comm)
        t1.record()
        torch.cuda.synchronize()
    ms = torch.tensor([t0.elapsed_time(t1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    rmax = comm.all_reduce_sum(slab.eng.sumsq(slab.residual(y)))
    out["timing"] = {"nx": nx, "n_gpus": world, "iterations": int(its[0]), "ms_total": float(ms[0]),
                     "ms_per_newton_iteration": float(ms[0]) / max(int(its[0]), 1),
                     "residual_l2_after": float(torch.sqrt(rmax)[0])}
    return out


def amr_on_slabs(rank, world, dev, n=1_000_000):
    """Refiner on a slab-decomposed mesh: three fronts on n cells, pruning on; the gathered result
    must equal the single-GPU Refiner's mesh bit for bit.  Reported, never raised on one rank only."""
    import torch
    import torch.distributed as dist
    from splitfxm_b200.distributed import slab_bounds
    from splitfxm_b200.refine import Refiner
    from splitfxm_b200.slab_refine import SlabRefiner
    from splitfxm_b200.spike import DistComm, ThreadGroup
    comm = DistComm() if world > 1 else ThreadGroup(1).comm(0)
    x = np.linspace(0.0, 1.0, n)
    U = np.stack([np.tanh(4000 * (x - 0.3)), np.where(x < 0.5, 1.0, 0.125) + 0.05 * np.sin(40 * x),
                  np.exp(-((x - 0.7) / 0.0005) ** 2)], axis=1)
    crit = (2.0, 0.05, 0.05, 0.01)
    lo, hi = slab_bounds(n, world, rank)
    r = SlabRefiner()
    r.set_criteria(*crit)
    r.set_max_points(4 * n)
    xd, Ud = torch.from_numpy(x[lo:hi].copy()).to(dev), torch.from_numpy(U[lo:hi].copy()).to(dev)
    for timed in (False, True):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record()
        x_new, U_new, lo_new, total = r.refine_slab(xd, Ud, lo, n, comm, rebalance=True, report=False)
        t1.record()
        torch.cuda.synchronize()
    ms = torch.tensor([t0.elapsed_time(t1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    mine = (lo_new, x_new.cpu().numpy(), U_new.cpu().numpy())
    parts = [mine]
    if world > 1:
        parts = [None] * world
        dist.all_gather_object(parts, mine)
    if rank != 0:
        return None
    parts.sort(key=lambda p: p[0])
    ref = Refiner()
    ref.set_criteria(*crit)
    ref.set_max_points(4 * n)
    x_ref, U_ref, _ = ref.refine_arrays(x, U, want_flags=False)
    same = bool(np.array_equal(np.concatenate([p[1] for p in parts]), x_ref)
                and np.array_equal(np.concatenate([p[2] for p in parts]), U_ref))
    return {"cells_before": n, "cells_after": int(total), "slab_sizes_after": [int(p[1].shape[0]) for p in parts],
            "ms_refine_incl_rebalance": float(ms[0]), "identical_to_single_gpu": same}


def main():
    if os.environ.get("SFXM_HANG_DUMP"):          # debugging aid: print where every rank is after N seconds and exit
        import faulthandler
        faulthandler.dump_traceback_later(int(os.environ["SFXM_HANG_DUMP"]), exit=True)
    ap = argparse.ArgumentParser()
    ap.add_argument("--nx", type=int, default=10_000_000)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--check", action="store_true")
    ap.add_argument("--fused", action="store_true", help="halo exchange fused into the update kernel (peer memory)")
    ap.add_argument("--rk45-check", action="store_true", help="also check adaptive RK45 on slabs against one GPU")
    ap.add_argument("--newton", action="store_true",
                    help="also run Newton on slabs (partitioned block solve): parity check + timing at --newton-nx")
    ap.add_argument("--newton-nx", type=int, default=1_000_000)
    ap.add_argument("--amr", action="store_true", help="also refine a slab-decomposed mesh and compare with one GPU")
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    import splitfxm_b200 as sf
    from splitfxm_b200.distributed import SlabSimulation
    from splitfxm_b200.models.euler1d import Euler1DConservative
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device(f"cuda:{local}")
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    def make(nx):
        x = np.linspace(0.0, 1.0, nx)
        slab = SlabSimulation(x, COMPS, "euler_cons", {"gamma": 1.4}, BCS, sf.FVSchemes.LAX_FRIEDRICHS, "FVM", rank, world,
                              device=dev)
        y = slab.eng.to_device(sod_state(x[slab.lo:slab.hi]).reshape(-1))
        return x, slab, y

    check = None
    if args.check:
        nx = 4096
        x, slab, y = make(nx)
        dt = 0.4 * (x[1] - x[0]) / 2.2
        if args.fused and world > 1:
            slab.enable_peer_halos()
            y, nsteps = slab.evolve_euler_fused(y, 40 * dt, dt)
        else:
            y, nsteps = slab.evolve_euler(y, 40 * dt, dt)
        mine = y.cpu().numpy().reshape(-1, 3)
        if world > 1:
            parts = [None] * world
            dist.all_gather_object(parts, mine)
        else:
            parts = [mine]
        if rank == 0:
            d = sf.Domain.from_size(nx, 1, 1, COMPS)
            d.interior_U()[:, :] = sod_state(np.array(d.positions(interior=True)))
            sim = sf.Simulation(d, Euler1DConservative(1.4, "FVM"), {}, BCS, sf.FVSchemes.LAX_FRIEDRICHS)
            sim.evolve(40 * dt, method="Euler", max_step=dt)
            check = bool(np.array_equal(np.concatenate(parts), d.values(interior=True)))   # reported, never raised:
            # an exception on one rank would leave the others waiting in the next collective

    rk_check = None
    if args.rk45_check:
        nx = 3000
        x, slab, y = make(nx)
        yk, nfev = slab.evolve_rk45(y, 0.02)
        mine = yk.cpu().numpy()[:, 0]
        if world > 1:
            parts = [None] * world
            dist.all_gather_object(parts, (slab.lo, slab.hi, mine))
        else:
            parts = [(slab.lo, slab.hi, mine)]
        if rank == 0:
            U = np.concatenate([v.reshape(hi - lo, 3) for lo, hi, v in parts])
            d = sf.Domain.from_size(nx, 1, 1, COMPS)
            d.interior_U()[:, :] = sod_state(np.array(d.positions(interior=True)))
            sim = sf.Simulation(d, Euler1DConservative(1.4, "FVM"), {}, BCS, sf.FVSchemes.LAX_FRIEDRICHS)
            sim.evolve(0.02)
            ref = d.values(interior=True)
            err = float(np.max(np.abs(U - ref)) / np.max(np.abs(ref)))
            # the slabs sum the error norm in a different order than one GPU does: step sizes agree to round-off only
            rk_check = {"nfev": nfev, "max_rel_diff_vs_single_gpu": err, "ok": bool(err < 1e-9)}

    amr = amr_on_slabs(rank, world, dev) if args.amr else None
    newton = None
    if args.newton:
        newton = newton_on_slabs(args, rank, world, dev)

    x, slab, y = make(args.nx)
    dt = 0.4 * (x[1] - x[0]) / 2.2
    other = torch.empty_like(y)

    def step(y, other):
        slab.exchange(y)
        slab.eng.euler_step(y, other, dt)
        return other, y

    fused = args.fused and world > 1
    if fused:
        slab.enable_peer_halos()
        wsteps = max(4, args.warmup + args.warmup % 2)     # even: the timed call then starts from the same pair of vectors
        y, _ = slab.evolve_euler_fused(y, wsteps * dt * 0.999999, dt)   # and replays the step pair captured here
    else:
        for _ in range(args.warmup):
            y, other = step(y, other)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    if fused:
        y, nst = slab.evolve_euler_fused(y, args.steps * dt * 0.999999, dt)     # exactly args.steps steps
        assert nst == args.steps, nst
    else:
        for _ in range(args.steps):
            y, other = step(y, other)
    t1.record()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ms = torch.tensor([t0.elapsed_time(t1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms = float(ms[0])
    finite = bool(torch.isfinite(y).all())
    if rank == 0:
        per_step = ms / args.steps
        cells_s = args.nx * args.steps / (ms * 1e-3)
        print(json.dumps({"workload": f"conservative Sod, Euler1DConservative FVM LF, nx={args.nx}, explicit Euler, "
                                      f"slab-decomposed x{world} with " + ("halo exchange fused into the update kernel (peer stores over NVLink)" if (args.fused and world > 1) else "NCCL halo exchange"),
                          "n_gpus": world, "steps": args.steps, "ms_per_step": per_step, "cell_updates_per_s": cells_s,
                          "GBps_per_gpu": (args.nx / world) * 56 / (per_step * 1e-3) / 1e9, "scaling": "strong",
                          "parity_vs_single_gpu": check, "rk45_on_slabs": rk_check, "newton_on_slabs": newton, "amr_on_slabs": amr,
                          "finite": finite}), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
