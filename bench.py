#!/usr/bin/env python
"""bench.py - headline benchmark of the B200-native SplitFXM hot path.

Workload (BASELINE.json configs[1], batched as configs[3]): finite-difference sparse-Jacobian assembly of the
nx = 250, 3-variable advection-diffusion problem (benchmark/test_jac_benchmark.py:14-61 shape, FDM CENTRAL,
1+1 ghost cells, split [1]) for an ensemble of `--members` independent (c, nu) members per GPU.
One step = one assembly of every member's block-banded Jacobian (one kernel launch).

  value      assemblies/s, inputs resident in HBM, CUDA-event time, max over ranks
  e2e        same metric through EnsembleSimulation.jacobian_host with HOST buffers (pinned H2D of the states and
             D2H of the banded Jacobians inside the timed region); e2e.pcie_probe: what the host link gives when all
             ranks copy at once
  roofline   the Jacobian kernel: bytes actually moved / its CUDA-event duration vs the measured HBM peak (`frac`),
             the SURVEY 8(d) figure beside it, and the fraction of the FP64 pipe, which is the nearer bound
  newton     BASELINE metric, second half: Split-Newton steady states/s of the same ensemble, on every rank
             (value device-resident, e2e through host buffers)
  slab       N > 1 only: BASELINE config 5 (one conservative Sod mesh of 1e7 cells slab-decomposed over the ranks,
             halo exchange fused into the update kernel vs NCCL) and Newton on slabs - the collective paths
  cpu_baseline  the UNMODIFIED reference (baseline/_ref) on ONE host core, bounded sample of the same members
                (kind "reference"; the oracle port only if baseline/_ref is not installed)
  extra      secondary rooflines (ensemble residual / solves, one-mesh kernels, AMR, RK45) - rank 0 only

`--impl reference` times the unmodified reference's own `Simulation.jacobian` / `Simulation.steady_state` on all host
cores (one member per core and step), same `config`, `metric`, `unit`.  Timed regions run with the cyclic GC off.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

NX, NV, W = 250, 3, 3
COMPS = ["u", "v", "w"]
ICS = {"u": "gaussian", "v": "rarefaction", "w": "tophat"}
# benchmark BCs as main.py:14-27 except u, which is dirichlet so that every block is non-singular and the same
# ensemble can also be Newton-solved
BCS = {"u": {"left": {"dirichlet": 0.0}, "right": {"dirichlet": 0.0}},
       "v": {"left": {"dirichlet": 3}, "right": {"dirichlet": 4}},
       "w": {"left": {"dirichlet": 2}, "right": "periodic"}}
SPLIT, LOCS = True, [1]
METRIC = "jacobian_assemblies_per_s"
UNIT = "assemblies/s (nx=250, nv=3, FDM CENTRAL, split [1])"
NEWTON_METRIC = "newton_solves_per_s"
NEWTON_UNIT = "steady states/s (Split-Newton to convergence, nx=250, nv=3, FDM CENTRAL, split [1])"


def member_params(first, count, total):
    """(c, nu) of members [first, first+count) of a sqrt(total) x sqrt(total) sweep grid."""
    side = max(int(round(total ** 0.5)), 1)
    k = np.arange(first, first + count)
    c = np.linspace(0.05, 0.5, side)[(k // side) % side]
    nu = np.logspace(-4, -2, side)[k % side]
    return c, nu


def initial_state():
    """Flat split-major initial vector [750] of the benchmark problem, built with the product's host API
    (the GPU arm never touches oracle/)."""
    import splitfxm_b200 as sf
    from splitfxm_b200.initialize import set_initial_condition
    d = sf.Domain.from_size(NX, 1, 1, COMPS)
    for c in COMPS:
        set_initial_condition(d, c, ICS[c])
    return np.array(d.positions(interior=True)), d.listify_interior(SPLIT, LOCS)


def initial_state_oracle():
    """The same vector from the oracle's own restatement (CPU arm)."""
    from oracle import core
    x = np.linspace(0.0, 1.0, NX)
    U = np.stack([core.initial_condition(x, ICS[c]) for c in COMPS], axis=1)
    return x, core.to_flat(U, SPLIT, LOCS)


def member_states(y0, first, count):
    """Deterministic per-member perturbation of the initial state (no RNG: a smooth function of member id)."""
    k = np.arange(first, first + count, dtype=np.float64)
    q = np.arange(y0.shape[0], dtype=np.float64)
    return y0[:, None] + 1e-3 * np.sin(0.37 * q[:, None] + 0.011 * k[None, :])


# ------------------------------------------------------------------------------------------------
# CPU arms: the unmodified reference (baseline/_ref) when it is installed, else the oracle port
# ------------------------------------------------------------------------------------------------
def bench_config(members_per_gpu, world):
    """`config` of the JSON line - the same object on the GPU arm and on the reference arm."""
    total = members_per_gpu * world
    return {"workload": f"configs[1] sparse-Jacobian benchmark shape (nx={NX}, nv={NV}, FDM CENTRAL, nb 1/1, split {LOCS}) "
                        f"for an ensemble of {members_per_gpu} members per GPU (configs[3] sweep over c, nu; {total} members "
                        "in total); Newton leg: Split-Newton steady state of every member of the same ensemble",
            "members_per_gpu": members_per_gpu, "parallelism": f"ensemble-sharded x{world}, no collective",
            "l2": "inputs+outputs per launch (1.6 GB moved) exceed the 126 MB L2"}


def _cpu_assemble(args):
    first, count, total = args
    from oracle import core
    x, y0 = initial_state_oracle()
    c, nu = member_params(first, count, total)
    Y = member_states(y0, first, count)
    acc = 0.0
    for k in range(count):
        P = core.Problem(core.AdvectionDiffusion(c[k], nu[k]), "FDM", "CENTRAL", x, 1, 1, BCS, COMPS)
        Jb = P.jacobian_banded(core.to_block(Y[:, k], NX, NV, SPLIT, LOCS))
        acc += float(Jb[0, 1, 0, 0])
    return acc


def reference_available():
    from baseline import reference_env
    return reference_env.available()


def cpu_baseline(sample, newton_sample, total):
    """Reported baseline on ONE host core, bounded sample of the same ensemble: the unmodified reference
    (`kind: reference`, baseline/_ref, in a spawned process) or, if it is not installed, the oracle port."""
    cores_box = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    if reference_available():
        from baseline.ref_worker import RefPool
        n_j = max(sample // 160, 8)                 # the reference is ~150x slower than the port: 64 assemblies ~ 11 s
        pool = RefPool(1, n_j, total)
        try:
            dt, _ = pool.step("assemble")
        finally:
            pool.close()
        out = {"value": n_j / dt, "unit": UNIT, "cores": 1, "kind": "reference",
               "sample": f"{n_j} members of the same ensemble, unmodified reference Simulation.jacobian(x0, True, [1]) "
                         f"(simulation.py:270-394, baseline/_ref), {dt:.2f} s on 1 of {cores_box} host cores"}
        if newton_sample > 0:
            pool = RefPool(1, newton_sample, total)
            try:
                dtn, res = pool.step("steady")
            finally:
                pool.close()
            its = res[0][1]
            out["newton"] = {"value": newton_sample / dtn, "unit": NEWTON_UNIT, "cores": 1, "kind": "reference",
                             "iterations_min_max": [min(its), max(its)],
                             "sample": f"{newton_sample} members, reference Simulation.steady_state(split=True, split_locs=[1]) "
                                       f"over the oracle/newton.py stand-in for the absent splitnewton, {dtn:.2f} s on 1 core"}
        return out
    t0 = time.perf_counter()
    _cpu_assemble((0, sample, total))
    dt = time.perf_counter() - t0
    return {"value": sample / dt, "unit": UNIT, "cores": 1, "kind": "port",
            "sample": f"{sample} members of the same ensemble, oracle/core.py Problem.jacobian_banded (NumPy colour "
                      f"passes), {dt:.2f} s on 1 core (baseline/_ref not installed)"}


def run_reference(args):
    """Reference arm: the reference's own CPU implementation on all host cores, bounded sample per step.  Under
    torchrun only rank 0 works."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    total = args.members * args.gpus
    cfg = bench_config(args.members, args.gpus)
    newton = None
    if reference_available():
        from baseline.ref_worker import RefPool
        kind, per_proc = "reference", max(args.ref_members_per_core, 1)
        sample = per_proc * cores
        pool = RefPool(cores, per_proc, total)
        try:
            for _ in range(args.warmup):
                pool.step("assemble")
            dt = 0.0
            for _ in range(args.steps):
                dt += pool.step("assemble")[0]
            n_steps = max(1, min(args.steps, args.newton_steps))
            pool.step("steady")
            dtn, its = 0.0, []
            for _ in range(n_steps):
                t, res = pool.step("steady")
                dtn += t
                its = [i for r in res for i in r[1]]
        finally:
            pool.close()
        what = (f"{sample} members per step ({per_proc} per process over {cores} processes), unmodified reference "
                "Simulation.jacobian (baseline/_ref)")
        newton = {"metric": NEWTON_METRIC, "value": sample * n_steps / dtn, "unit": NEWTON_UNIT, "steps": n_steps,
                  "ms_per_step": 1e3 * dtn / n_steps, "iterations_min_max": [min(its), max(its)],
                  "sample": f"{sample} members per step, reference Simulation.steady_state(split=True, split_locs=[1]) over "
                            "the oracle/newton.py stand-in for the absent third-party splitnewton (parity unpinned)"}
    else:
        import multiprocessing as mp
        kind = "port"
        per_proc = max(args.ref_members // cores, 1)
        sample = per_proc * cores
        jobs = [(i * per_proc, per_proc, total) for i in range(cores)]
        with mp.get_context("fork").Pool(cores) as pool:
            for _ in range(args.warmup):
                pool.map(_cpu_assemble, jobs)
            t0 = time.perf_counter()
            for _ in range(args.steps):
                pool.map(_cpu_assemble, jobs)
            dt = time.perf_counter() - t0
        what = f"{sample} members per step over {cores} processes (oracle/core.py; baseline/_ref not installed)"
    value = sample * args.steps / dt
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": cfg,
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind, "sample": what},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "newton": newton}
    print(json.dumps(line), flush=True)


class quiet_host:
    """Timed regions run with Python's cyclic garbage collector off (as `timeit` does): a generation-2 collection of an
    interpreter that has imported torch + scipy pauses the host for tens of milliseconds, which showed up as 5.9 / 7.7 /
    27 ms outliers among 4.4 ms Newton steps (the legs whose kernels are launched step by step from Python)."""

    def __enter__(self):
        import gc
        gc.collect()
        self.was = gc.isenabled()
        gc.disable()
        return self

    def __exit__(self, *exc):
        import gc
        if self.was:
            gc.enable()
        return False


# ------------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    """SM clock and throttle reasons DURING the timed region: NVML polled every millisecond from a thread (an
    nvidia-smi process takes longer to start than the timed region lasts); falls back to nvidia-smi."""
    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.stop_flag = index, threading.Event()
        self.sm, self.mx, self.reasons = [], None, set()
        self.nvml = None
        try:
            import pynvml
            pynvml.nvmlInit()
            # NVML indexes physical devices; honour CUDA_VISIBLE_DEVICES if it is a plain index list
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = index
            if vis:
                ids = [v.strip() for v in vis.split(",") if v.strip()]
                if index < len(ids) and ids[index].isdigit():
                    phys = int(ids[index])
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.mx = float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
            self.nvml = pynvml
        except Exception:
            self.nvml = None

    def _sample(self):
        if self.nvml is not None:
            n = self.nvml
            self.sm.append(float(n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM)))
            try:
                mask = n.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle)
            except Exception:
                mask = 0
            for bit, name in self.REASONS.items():
                if mask & bit:
                    self.reasons.add(name)
            return
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        out = subprocess.run(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-i", str(self.index)],
                             capture_output=True, text=True, timeout=5).stdout
        r = [f.strip() for f in out.strip().split(",")]
        if len(r) >= 6:
            self.sm.append(float(r[0]))
            self.mx = float(r[1])
            for name, v in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], r[2:6]):
                if v.lower().startswith("active"):
                    self.reasons.add(name)

    def run(self):
        while not self.stop_flag.is_set():
            try:
                self._sample()
            except Exception:
                pass
            self.stop_flag.wait(0.001 if self.nvml is not None else 0.1)

    def summary(self):
        return {"sm_mhz": float(np.median(self.sm)) if self.sm else None, "sm_max_mhz": self.mx,
                "reasons": sorted(self.reasons), "samples": len(self.sm),
                "source": "nvml" if self.nvml is not None else "nvidia-smi"}


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
def run_gpu(args):
    import torch
    import torch.distributed as dist
    import splitfxm_b200 as sf
    from splitfxm_b200 import _lib
    from splitfxm_b200.ensemble import EnsembleSimulation

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch with torch.distributed.run --nproc-per-node N for --gpus N")
    torch.cuda.set_device(local)
    if world > 1 or os.environ.get("SFXM_BENCH_FORCE_DIST"):      # (the env switch: NCCL initialised on one GPU, for diagnosis)
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29533")
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"), rank=rank, world_size=world)
    dev = torch.device(f"cuda:{local}")
    B = args.members
    total = B * world
    x, y0 = initial_state()
    c, nu = member_params(rank * B, B, total)          # members shard across ranks: no data-path collective
    d = sf.Domain.from_size(NX, 1, 1, COMPS)
    ens = EnsembleSimulation(d, "ad", {"c": c, "nu": nu}, BCS, sf.FDSchemes.CENTRAL, method="FDM", device=dev)
    eng = ens.eng
    Y_host = torch.from_numpy(member_states(y0, rank * B, B)).pin_memory()
    Y = Y_host.to(dev, non_blocking=True)
    J = eng.jacobian_buffer()
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident timing --------------------------------------------------------------------
    for _ in range(args.warmup):
        eng.jacobian(Y, SPLIT, LOCS, J=J)
    sampler = ClockSampler(local) if rank == 0 else None
    barrier()
    if sampler:
        sampler.start()
    _lib.LAUNCHES[0] = 0
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    with quiet_host():
        ev[0].record()
        for k in range(args.steps):
            eng.jacobian(Y, SPLIT, LOCS, J=J)
            ev[k + 1].record()
        barrier()
    launches = _lib.LAUNCHES[0]
    if sampler:
        sampler.stop_flag.set()
        sampler.join()
    elapsed_ms = ev[0].elapsed_time(ev[-1])
    per_launch_ms = [ev[k].elapsed_time(ev[k + 1]) for k in range(args.steps)]
    t = torch.tensor([elapsed_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    elapsed_ms = float(t[0])
    value = total * args.steps / (elapsed_ms * 1e-3)

    # ---- roofline of the Jacobian kernel (this rank's launches) ------------------------------------
    # `frac` is on the bytes the kernel actually moves: read U (nv*8 per cell-member) + x (8 per cell, shared by the
    # members), write W*nv entries per cell-member.  Advection-diffusion is one scalar equation per variable, so the
    # off-diagonal same-segment entries of a block are exact zeros: structural - zero-filled once by the buffer's owner,
    # never re-written - 24 + W*NV*8 = 96 B per cell-member; ncu dram__bytes (`traffic`) = 1.01x that.  SURVEY 8(d)'s
    # per-unit figure counts all same-segment entries (24 + W * sum s_k^2 * 8 = 144 B): kept as `survey_8d_*`.
    seg_entries = sum(s * s for s in (LOCS[0], NV - LOCS[0]))
    survey_bytes = B * NX * (NV * 8 + W * seg_entries * 8) + NX * 8
    moved_bytes = B * NX * (NV * 8 + W * NV * 8) + NX * 8
    kern_ms = float(np.mean(per_launch_ms))
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = json.load(open(peaks_path))["hbm_gbs"], "measured (MEASURED_PEAKS.json hbm_gbs)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    achieved = moved_bytes / (kern_ms * 1e-3) / 1e9
    traffic, fp64_instr = None, None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        tj = json.load(open(tpath))
        traffic = tj.get("jacobian_kernel_dram_bytes_per_launch")
        fp64_instr = tj.get("jacobian_kernel_fp64_instr_per_cell_member")
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "kernel": JAC_KERNEL,
                "algorithmic_bytes_per_launch": moved_bytes, "kernel_ms": kern_ms, "peak_source": peak_src,
                "survey_8d_bytes_per_launch": survey_bytes, "survey_8d_frac": survey_bytes / (kern_ms * 1e-3) / 1e9 / peak,
                "note": "frac = bytes actually moved / kernel time / measured HBM peak; survey_8d_* counts the structural zeros "
                        "that are never stored (SURVEY 8(d)'s per-unit figure), so it can exceed 1"}
    if fp64_instr:
        # FP64 pipe: 64 lanes per SM per clock x 148 SMs at the sampled SM clock
        rate = 148 * 64 * 1.965e9
        roofline["fp64_pipe_frac"] = fp64_instr * B * NX / rate / (kern_ms * 1e-3)
        roofline["fp64_instr_per_cell_member"] = fp64_instr
        roofline["note"] += ("; fp64_pipe_frac = FP64-pipe instructions (ncu) x cell-members / (148 SMs x 64 lanes x "
                             "1.965 GHz) / kernel time - since the three-operation exact division the FP64 pipe needs about as long "
                             "as the bytes do at the HBM peak (0.22 vs 0.24 ms)")

    # ---- Newton leg: Split-Newton steady state of every member, on every rank ----------------------
    newton = newton_leg(args, torch, dist, ens, Y_host, dev, world, total, barrier, peak)

    # ---- end to end through the public API with host buffers --------------------------------------
    # Every step uploads its states from pinned host memory and downloads its Jacobians (the same-segment entries -
    # the only ones that exist in split mode) into pinned host memory.  Two slots: the download of step k overlaps
    # the upload + assembly of step k+1; every step's transfers are inside the timed region.
    e2e_steps = max(2, min(args.steps, args.e2e_steps))
    n_live = len(eng.live_entries(SPLIT, LOCS))
    Jh = [torch.empty((NX, W, n_live, B), dtype=torch.float64).pin_memory() for _ in range(2)]
    for k in range(2):
        ens.jacobian_host(Y_host, Jh[k], SPLIT, LOCS, slot=k)
    ens.sync_host()
    barrier()
    with quiet_host():
        t0 = torch.cuda.Event(enable_timing=True)
        t0.record()
        for k in range(e2e_steps):
            ens.jacobian_host(Y_host, Jh[k % 2], SPLIT, LOCS, slot=k % 2)
        ens.sync_host()
        t1 = torch.cuda.Event(enable_timing=True)
        t1.record()
        barrier()
    torch.cuda.synchronize()
    te = torch.tensor([t0.elapsed_time(t1)], dtype=torch.float64, device=dev)   # t1 is recorded after both streams drained
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e = {"value": total * e2e_steps / (float(te[0]) * 1e-3), "unit": UNIT,
           "h2d_bytes_per_step": int(Y_host.numel() * 8), "d2h_bytes_per_step": int(Jh[0].numel() * 8),
           "steps": e2e_steps,
           "api": "EnsembleSimulation.jacobian_host(pinned host states -> pinned host Jacobians), 2 slots: the download "
                  "of a step overlaps the upload + assembly of the next; only the entries the reference's sparse "
                  f"matrix holds ({n_live} of {NV * NV} per block: same-segment entries, minus the exact zeros of the "
                  "decoupled advection-diffusion equations)"}
    checksum = float(Jh[(e2e_steps - 1) % 2][:, 1, 0, :].sum())
    e2e["checksum_diag_u"] = checksum
    # what the host link gives when every rank copies at once: e2e is bound by the device->host copy of the Jacobians,
    # so its scaling over the GPUs of one box is the scaling of this number (measured, not assumed)
    probe = torch.empty(32 * 1024 * 1024, dtype=torch.float64, device=dev)          # 256 MB
    host = Jh[0].reshape(-1)[:probe.numel()] if Jh[0].numel() >= probe.numel() else torch.empty(probe.numel(), dtype=torch.float64).pin_memory()
    rates = []
    for direction in ("d2h", "h2d"):
        barrier()
        p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        p0.record()
        for _ in range(4):
            if direction == "d2h":
                host.copy_(probe, non_blocking=True)
            else:
                probe.copy_(host, non_blocking=True)
        p1.record()
        torch.cuda.synchronize()
        rates.append(4 * probe.numel() * 8 / (p0.elapsed_time(p1) * 1e-3) / 1e9)
    rt = torch.tensor(rates, dtype=torch.float64, device=dev)
    if world > 1:
        allr = [torch.empty_like(rt) for _ in range(world)]
        dist.all_gather(allr, rt)
        allr = torch.stack(allr).cpu().numpy()
    else:
        allr = rt.cpu().numpy()[None, :]
    e2e["pcie_probe"] = {"what": "all ranks copy 4 x 256 MB pinned <-> device at the same time (GB/s per rank)",
                         "d2h_GBps_per_rank": [round(float(v), 1) for v in allr[:, 0]],
                         "h2d_GBps_per_rank": [round(float(v), 1) for v in allr[:, 1]],
                         "d2h_GBps_aggregate": round(float(allr[:, 0].sum()), 1),
                         "e2e_d2h_GBps_aggregate": round(e2e["value"] / total * world * e2e["d2h_bytes_per_step"] / 1e9, 1)}
    del Jh, probe, host

    # ---- one large mesh, slab-decomposed over the ranks (BASELINE config 5): every rank, N > 1 only ----
    slab = None
    if world > 1 and not args.no_slab:
        del ens._slots
        torch.cuda.empty_cache()
        slab = slab_leg(args, torch, dist, sf, dev, rank, world, peak)

    # ---- extras: Newton iterations on the same ensemble, single-mesh sweep -------------------------
    extra = {}
    if rank == 0 and not args.no_extra:
        extra = extras(torch, sf, ens, Y, dev, peak)

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": elapsed_ms / args.steps, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": bench_config(B, world),
                "vs_baseline_note": "BASELINE.json publishes nothing for this metric (published = {}); the reference's README "
                                    "quotes ~0.6 s per sparse assembly at N=250 on unstated hardware (README.md:100-105), i.e. "
                                    "%.3g x that figure - not a like-for-like baseline, hence null" % (value * 0.6),
                "gpu_launches": launches, "roofline": roofline, "e2e": e2e, "newton": newton, "slab": slab,
                "cpu_baseline": cpu_baseline(args.cpu_sample, args.cpu_newton_sample, total),
                "clocks": sampler.summary() if sampler else None, "extra": extra}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


JAC_KERNEL = "jacobian_pair_kernel<ModelAD<3>, FD_CENTRAL, 3, 3, false>"


def newton_leg(args, torch, dist, ens, Y_host, dev, world, total, barrier, peak):
    """BASELINE metric, second half: Newton solves/s = Split-Newton steady states of the ensemble per second.
    `value`: states resident in HBM; `e2e`: EnsembleSimulation.steady_state_host, pinned host states in, pinned host
    steady states + iteration counts out, copies inside the timed region.  Runs on every rank; max over ranks."""
    from splitfxm_b200 import _lib
    eng = ens.eng
    B = eng.batch
    Y = Y_host.to(dev)
    steps = max(1, min(args.steps, args.newton_steps))
    for _ in range(3):       # warm-up with the timed loop's own assignment pattern: the caching allocator reaches its steady
        Yf, its = ens.steady_state(Y, split=SPLIT, split_locs=LOCS)      # state (a first-time cudaMalloc of a 393 MB state
    barrier()                                                            # block showed up as one 31 ms step among 4.3 ms ones)
    _lib.LAUNCHES[0] = 0
    marks = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
    host_s = []
    with quiet_host():
        marks[0].record()
        for k in range(steps):
            h0 = time.perf_counter()
            Yf, its = ens.steady_state(Y, split=SPLIT, split_locs=LOCS)
            host_s.append(time.perf_counter() - h0)
            marks[k + 1].record()
        t0, t1 = marks[0], marks[-1]
        barrier()
    launches = _lib.LAUNCHES[0]
    per_step = [marks[k].elapsed_time(marks[k + 1]) for k in range(steps)]
    t = torch.tensor([t0.elapsed_time(t1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t[0]) / steps
    res = float(torch.max(torch.abs(eng.residual(Yf, None, SPLIT, LOCS))))
    it_lo, it_hi = int(its.min()), int(its.max())
    # end to end: two slots (device state buffers + pinned result buffers); the download of a step overlaps the upload and
    # the solve of the next one, and the next step's upload runs while this step's Newton iterations do
    outs = [torch.empty((eng.n, B), dtype=torch.float64).pin_memory() for _ in range(2)]
    itss = [torch.empty(B, dtype=torch.int64).pin_memory() for _ in range(2)]
    out_host, its_host = outs[0], itss[0]
    esteps = max(steps, 12)

    def e2e_pass(n):
        for k in range(n):
            ens.steady_state_host(Y_host, outs[k % 2], itss[k % 2], SPLIT, LOCS, slot=k % 2,
                                  next_Y_host=Y_host if k + 1 < n else None)
        ens.sync_host()

    e2e_pass(3)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with quiet_host():
        e0.record()
        w0 = time.perf_counter()
        e2e_pass(esteps)
        wall_ms = (time.perf_counter() - w0) * 1e3
        e1.record()
        barrier()
    # the pass ends with sync_host() (all three streams drained): wall clock and the event pair bracket the same work;
    # take the larger of the two, then the max over ranks
    te = torch.tensor([max(e0.elapsed_time(e1), wall_ms)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    ms_e = float(te[0]) / esteps
    # linear-solve sweeps actually run: per outer iteration the inner segment is solved to convergence, then one step
    # on the outer segment; fused algorithmic bytes per cell-member per sweep: s = 1 -> 48, s = 2 -> 128
    return {"metric": NEWTON_METRIC, "value": total / (ms * 1e-3), "unit": NEWTON_UNIT, "steps": steps, "ms_per_step": ms,
            "gpu_launches": launches, "iterations_min_max": [it_lo, it_hi], "max_abs_residual": res,
            "per_step_ms_rank0": [round(v, 3) for v in per_step], "host_ms_per_call_rank0": [round(1e3 * v, 3) for v in host_s],
            "e2e": {"value": total / (ms_e * 1e-3), "unit": NEWTON_UNIT, "ms_per_step": ms_e,
                    "h2d_bytes_per_step": int(Y_host.numel() * 8), "d2h_bytes_per_step": int(out_host.numel() * 8 + B * 8),
                    "steps": esteps,
                    "api": "EnsembleSimulation.steady_state_host(pinned host states -> pinned host steady states + "
                           "iteration counts), 2 slots over 3 streams: the download of a step overlaps the upload and the "
                           "solve of the next, the next upload runs during this step's Newton iterations",
                    "checksum": float(out_host[:, 0].sum())}}


def slab_leg(args, torch, dist, sf, dev, rank, world, peak):
    """BASELINE config 5 under the driver: ONE conservative Sod mesh (Euler1DConservative, FVM Lax-Friedrichs, neumann
    ends), nx = --slab-nx cells in total, slab-decomposed over the ranks; explicit Euler steps with the halo exchange
    (nb * nv doubles per side) fused into the update kernel as peer stores over NVLink, and the NCCL send/recv variant
    beside it.  Then Newton on a slab-decomposed advection-diffusion mesh (partitioned block solve, one small
    all-gather per solve), weak scaling: 1e6 cells per rank.  Nothing raises on one rank only."""
    from splitfxm_b200.distributed import SlabSimulation
    out = {}
    comps = ["rho", "rhou", "E"]
    bcs = {n: {"left": {"neumann": 0.0}, "right": {"neumann": 0.0}} for n in comps}

    def tmax(ms):
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t[0])

    try:
        nx = args.slab_nx
        x = np.linspace(0.0, 1.0, nx)
        slab = SlabSimulation(x, comps, "euler_cons", {"gamma": 1.4}, bcs, sf.FVSchemes.LAX_FRIEDRICHS, "FVM", rank, world,
                              device=dev)
        xs = x[slab.lo:slab.hi]
        left = xs < 0.5
        U0 = np.stack([np.where(left, 1.0, 0.125), np.zeros_like(xs), np.where(left, 2.5, 0.25)], axis=1)
        y = slab.eng.to_device(U0.reshape(-1))
        dt = 0.4 * (x[1] - x[0]) / 2.2
        steps = args.slab_steps
        slab.enable_peer_halos()
        y, _ = slab.evolve_euler_fused(y, 20 * dt * 0.999999, dt)
        dist.barrier()
        torch.cuda.synchronize()
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record()
        y, nst = slab.evolve_euler_fused(y, steps * dt * 0.999999, dt)
        t1.record()
        dist.barrier()
        torch.cuda.synchronize()
        ms = tmax(t0.elapsed_time(t1)) / max(nst, 1)
        other = torch.empty_like(y)
        for _ in range(5):
            slab.exchange(y)
            slab.eng.euler_step(y, other, dt)
            y, other = other, y
        dist.barrier()
        torch.cuda.synchronize()
        n0, n1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n0.record()
        for _ in range(max(steps // 4, 1)):
            slab.exchange(y)
            slab.eng.euler_step(y, other, dt)
            y, other = other, y
        n1.record()
        dist.barrier()
        torch.cuda.synchronize()
        ms_nccl = tmax(n0.elapsed_time(n1)) / max(steps // 4, 1)
        fin = torch.tensor([1.0 if bool(torch.isfinite(y).all()) else 0.0], dtype=torch.float64, device=dev)
        dist.all_reduce(fin, op=dist.ReduceOp.MIN)
        gbps = (nx / world) * 56 / (ms * 1e-3) / 1e9
        out["euler_sod"] = {"workload": f"conservative Sod, nx={nx}, explicit Euler (max_step = 0.4 dx / 2.2), slabs x{world}",
                            "scaling": "strong", "steps": int(nst), "ms_per_step_fused_peer_halo": ms,
                            "ms_per_step_nccl_halo": ms_nccl, "cell_updates_per_s": nx / (ms * 1e-3),
                            "GBps_per_gpu": gbps, "frac_of_hbm_peak_per_gpu": gbps / peak,
                            "bytes_per_cell_step": 56, "finite": bool(fin[0] > 0),
                            "collective": "none per step (peer stores + flag words over NVLink; the interior warps overlap the neighbour hand-shake); NCCL variant: "
                                          "batch_isend_irecv of nb*nv doubles per side"}
        del slab, y, other
        torch.cuda.empty_cache()
    except Exception as exc:                     # reported, never raised on one rank only
        out["euler_sod"] = {"error": repr(exc)}
    try:
        from splitfxm_b200.spike import DistComm
        ad_comps = ["u", "v", "w"]
        ad_bcs = {"u": {"left": {"dirichlet": 1.0}, "right": {"dirichlet": 0.0}},
                  "v": {"left": {"dirichlet": 3}, "right": {"dirichlet": 4}},
                  "w": {"left": {"dirichlet": 2}, "right": {"neumann": 0.0}}}
        nxn = args.slab_newton_nx * world
        xg = np.linspace(0.0, 1.0, nxn)
        slab = SlabSimulation(xg, ad_comps, "ad", {"c": 0.2, "nu": 0.05}, ad_bcs, sf.FVSchemes.LAX_FRIEDRICHS, "FVM", rank,
                              world, device=dev)
        xs = xg[slab.lo:slab.hi]
        U0 = np.stack([np.exp(-200 * (xs - 0.25) ** 2), np.where(xs > 0.5, 2.0, 0.0) * (xs - 0.5), (np.abs(xs - 0.5) < 0.25) * 1.0],
                      axis=1)
        y0 = slab.eng.to_device(U0.reshape(-1))
        comm = DistComm()
        for timed in (False, True):
            y = y0.clone()
            dist.barrier()
            torch.cuda.synchronize()
            t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0.record()
            its = slab.steady_state(y, False, None, comm=comm)
            t1.record()
            torch.cuda.synchronize()
        ms = tmax(t0.elapsed_time(t1))
        rn = comm.all_reduce_sum(slab.eng.sumsq(slab.residual(y)))
        out["newton_ad"] = {"workload": f"advection-diffusion FVM LF, nx={nxn} ({args.slab_newton_nx} cells per rank), unsplit "
                                        "Newton on slabs (partitioned block solve)", "scaling": "weak",
                            "iterations": int(its[0]), "ms_total": ms, "ms_per_newton_iteration": ms / max(int(its[0]), 1),
                            "newton_iterations_per_s": max(int(its[0]), 1) / (ms * 1e-3),
                            "residual_l2_after": float(torch.sqrt(rn)[0]),
                            "collective": "one all-gather of 2s(1+2s) doubles per rank per solve + all-reduced norms"}
        del slab, y, y0
        torch.cuda.empty_cache()
    except Exception as exc:
        out["newton_ad"] = {"error": repr(exc)}
    return out


def extras(torch, sf, ens, Y, dev, peak):
    """Secondary numbers (not the headline): Newton iteration throughput on the ensemble, residual and Jacobian
    on one large mesh."""
    from splitfxm_b200.models.advection_diffusion import AdvectionDiffusion
    from splitfxm_b200 import _lib as _lib_mod
    out = {}
    eng = ens.eng

    def timed(fn, reps):
        fn()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps):
            fn()
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) / reps

    B = eng.batch
    r = torch.empty_like(Y)
    s = torch.zeros_like(Y)
    J = eng.jacobian_buffer()
    status = torch.zeros(B, dtype=torch.int32, device=dev)
    ms_res = timed(lambda: eng.residual(Y, r, SPLIT, LOCS), 5)
    ms_solve = timed(lambda: eng.solve(J, r, s, SPLIT, LOCS, status=status), 3)
    def roof(kernel, nbytes, ms, what):
        return {"kernel": kernel, "bound": "hbm", "achieved": nbytes / ms / 1e6, "peak": peak, "unit": "GB/s",
                "frac": nbytes / ms / 1e6 / peak, "algorithmic_bytes_per_launch": int(nbytes), "kernel_ms": ms, "bytes": what}

    rooflines = []
    out["rooflines"] = rooflines
    rooflines.append(roof("residual_ring_kernel<ModelAD<3>, FD_CENTRAL> (ensemble residual)", B * NX * 48 + NX * 8, ms_res,
                          "read U 24 + write r 24 per cell-member"))
    out["ensemble_residual"] = {"ms": ms_res, "GBps": B * NX * 48 / ms_res / 1e6, "frac": B * NX * 48 / ms_res / 1e6 / peak}
    solve_bytes = B * NX * ((1 + 4) * W * 8 + 24 + 24)
    out["ensemble_block_tridiag_solve"] = {"ms": ms_solve, "GBps": solve_bytes / ms_solve / 1e6,
                                           "frac": solve_bytes / ms_solve / 1e6 / peak,
                                           "solves_per_s": B / (ms_solve * 1e-3)}
    # Newton linear step (both segments of split [1]): Jacobian assembled inside the Thomas sweep vs assemble-then-solve
    s2 = torch.zeros_like(Y)

    def fused_step():
        for k in range(2):
            eng.newton_solve_fused(Y, s2, SPLIT, LOCS, k, status=status)

    def unfused_step():
        for k in range(2):
            eng.jacobian(Y, SPLIT, LOCS, J=J, r0=r, seg_lo=k, seg_hi=k + 1)
            eng.solve(J, r, s, SPLIT, LOCS, k, k + 1, status=status, method="thomas")

    ms_f, ms_u = timed(fused_step, 3), timed(unfused_step, 3)
    # fused algorithmic bytes per cell-member: read the segment's U (8 s), write + read C' (8 s^2) and rhs' (8 s),
    # write the step (8 s): s = 1 -> 48, s = 2 -> 128
    fused_bytes = B * NX * (48 + 128)
    rooflines.append(roof("block_thomas_kernel<MemSrc> (ensemble block-tridiagonal solve, both segments)", solve_bytes, ms_solve,
                          "read J_seg 120 + f 24, write s 24 per cell-member"))
    rooflines.append(roof("block_thomas_kernel<FusedSrc> x2 (Newton linear step, Jacobian inside the sweep)", fused_bytes, ms_f,
                          "s=1: 48, s=2: 128 per cell-member (U, C', rhs', step)"))
    out["ensemble_newton_linear_step"] = {"fused_ms": ms_f, "assemble_then_solve_ms": ms_u,
                                          "bit_identical": bool(torch.equal(s, s2)),
                                          "linear_steps_per_s": B / (ms_f * 1e-3),
                                          "fused_GBps": fused_bytes / ms_f / 1e6, "fused_frac": fused_bytes / ms_f / 1e6 / peak}
    sweep = []
    for nx in (250, 1000, 10000, 100000, 1000000):
        d = sf.Domain.from_size(nx, 1, 1, COMPS)
        sim = sf.Simulation(d, AdvectionDiffusion(0.2, 0.001, "FDM"), ICS, BCS, sf.FDSchemes.CENTRAL)
        e1 = sim._engine()
        y = e1.to_device(d.listify_interior(SPLIT, LOCS))
        Jb = e1.jacobian_buffer()
        rr = torch.empty_like(y)
        ms_j = timed(lambda: e1.jacobian(y, SPLIT, LOCS, J=Jb), 5)
        ms_r = timed(lambda: e1.residual(y, rr, SPLIT, LOCS), 5)
        ss = torch.zeros_like(y)
        st1 = torch.zeros(1, dtype=torch.int32, device=dev)

        def newton_iteration():      # one Newton iteration of the device driver: J + f, solve (PCR), update + criterion norms
            e1.jacobian(y, SPLIT, LOCS, J=Jb, r0=rr)
            e1.solve(Jb, rr, ss, SPLIT, LOCS, status=st1)
            e1.update_norms(y, ss)
        ms_n = timed(newton_iteration, 3)
        U0 = d.values(interior=True)
        sim.steady_state(split=SPLIT, split_locs=LOCS)               # warm-up (allocations, first launches)
        d.interior_U()[:, :] = U0                                    # host write: the timed call uploads the state again
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        iters = sim.steady_state(split=SPLIT, split_locs=LOCS)       # public API: host state in, device-resident result,
        torch.cuda.synchronize()                                     # iteration count back
        t_ss = time.perf_counter() - t0
        if nx == 1000000:
            rooflines.append(roof("jacobian_scalar_kernel (one mesh, nx=1e6)", nx * (24 + 8 + 15 * 8), ms_j,
                                  "read U 24 + x 8, write 15 same-segment band entries per cell"))
            rooflines.append(roof("residual_scalar_kernel (one mesh, nx=1e6)", nx * 56, ms_r, "read U 24 + x 8, write r 24"))
            rooflines.append(roof("Newton iteration, one mesh nx=1e6 (Jacobian + nested-dissection PCR + update/norms)",
                                  nx * 640, ms_n, "SURVEY 8(d): 640 B per cell per Newton iteration"))
        sweep.append({"nx": nx, "jacobian_ms": ms_j, "jacobian_assemblies_per_s": 1e3 / ms_j,
                      "jacobian_GBps": nx * (24 + 8 + 15 * 8) / ms_j / 1e6, "residual_ms": ms_r,
                      "residual_GBps": nx * 56 / ms_r / 1e6, "newton_iteration_ms": ms_n,
                      "newton_iterations_per_s": 1e3 / ms_n, "steady_state_api_s": t_ss, "steady_state_iterations": iters})
    out["single_mesh_sweep"] = sweep

    # AMR pass and RK45 evolve on one large mesh (conservative Sod)
    from splitfxm_b200.models.euler1d import Euler1DConservative
    from splitfxm_b200.refine import Refiner
    import contextlib
    import io
    comps = ["rho", "rhou", "E"]
    nbc = {c: {"left": {"neumann": 0.0}, "right": {"neumann": 0.0}} for c in comps}
    for nx in (100000, 1000000, 10000000):
        d = sf.Domain.from_size(nx, 1, 1, comps)
        xs = d.interior_x()
        d.interior_U()[:, 0] = np.where(xs < 0.5, 1.0, 0.125) + 0.01 * np.tanh(200 * (xs - 0.3))
        d.interior_U()[:, 2] = np.where(xs < 0.5, 2.5, 0.25)
        r = Refiner()
        r.set_criteria(2.0, 0.05, 0.05, -0.001)
        r.set_max_points(4 * nx)
        xd = torch.from_numpy(np.ascontiguousarray(d.interior_x())).to(dev)
        Ud = torch.from_numpy(np.ascontiguousarray(d.interior_U())).to(dev)
        lib = _lib_mod.load()
        work = torch.empty(lib.sfxm_refine_fused_workspace_bytes(nx), dtype=torch.uint8, device=dev)
        counts = torch.zeros(3, dtype=torch.int64, device=dev)
        xo = torch.empty(2 * nx, dtype=torch.float64, device=dev)
        Uo = torch.empty((2 * nx, 3), dtype=torch.float64, device=dev)

        def amr_pass():      # statistics + marking + scan/compaction: three launches
            _lib_mod.check(lib.sfxm_refine_fused(xd.data_ptr(), Ud.data_ptr(), nx, 3, 2.0, 0.05, 0.05, -0.001, 0.01, 1e-10,
                                                 work.data_ptr(), xo.data_ptr(), Uo.data_ptr(), counts.data_ptr(),
                                                 _lib_mod.stream_ptr()))
        ms_amr = timed(amr_pass, 5)
        with contextlib.redirect_stdout(io.StringIO()):
            r.refine(sf.Domain.from_arrays(d._host()[0], d._host()[1], 1, 1, comps))      # warm-up
            d._device()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            r.refine(d)                                               # device-resident Domain in, device-resident out
            torch.cuda.synchronize()
        t_api = time.perf_counter() - t0
        if nx >= 1000000:
            rooflines.append(roof("AMR pass (sfxm_refine_fused: 3 launches), nx=%.0e" % nx, nx * 96, ms_amr,
                                  "SURVEY 8(d): ~96 B per cell over both passes (1e6 cells = 96 MB: launch-latency-bound)"))
        out.setdefault("amr", []).append({"nx": nx, "device_pass_ms": ms_amr, "GBps": nx * 96 / ms_amr / 1e6,
                                          "cells_after": int(counts[0]), "refine_api_s": t_api})
    # BASELINE config 3 in small: conservative Sod + AMR, the mesh growing from 60,000 cells past 1e5 (examples/sod_amr.py
    # runs the full recipe), every call through the reference-shaped API, the Domain staying on the device in between
    d = sf.Domain.from_size(60000, 1, 1, comps)
    xs = d.interior_x()
    d.interior_U()[:, 0] = np.where(xs < 0.5, 1.0, 0.125)
    d.interior_U()[:, 2] = np.where(xs < 0.5, 2.5, 0.25)
    sim = sf.Simulation(d, Euler1DConservative(1.4, "FVM"), {}, nbc, sf.FVSchemes.LAX_FRIEDRICHS)
    sim._r.set_criteria(2.0, 2e-5, 2e-5, -0.001)
    sim._r.set_max_points(400000)
    sim.evolve(2.0e-4, split=True, split_locs=[1])
    torch.cuda.synchronize()
    cells, stale = [d._nx], []
    before = _lib_mod.LAUNCHES[0]
    t0 = time.perf_counter()
    with contextlib.redirect_stdout(io.StringIO()):
        for _ in range(6):
            sim._r.refine(d)
            cells.append(d._nx)
            sim.evolve(1.0e-5, split=True, split_locs=[1])
            stale.append(not d._host_ok)
            if d._nx >= 100000:
                break
    torch.cuda.synchronize()
    t_c3 = time.perf_counter() - t0
    out["sod_amr_config3"] = {"cells": cells, "seconds": t_c3, "launches": _lib_mod.LAUNCHES[0] - before,
                              "domain_stayed_on_device": bool(all(stale)), "finite": bool(np.isfinite(d.values()).all())}

    d = sf.Domain.from_size(100000, 1, 1, comps)
    xs = d.interior_x()
    d.interior_U()[:, 0] = np.where(xs < 0.5, 1.0, 0.125)
    d.interior_U()[:, 2] = np.where(xs < 0.5, 2.5, 0.25)
    sim = sf.Simulation(d, Euler1DConservative(1.4, "FVM"), {}, nbc, sf.FVSchemes.LAX_FRIEDRICHS)
    U0 = d.values(interior=True)
    sim.evolve(2.0e-4, split=True, split_locs=[1])          # warm-up (allocations, tableau import)
    d.interior_U()[:, :] = U0
    torch.cuda.synchronize()
    before = _lib_mod.LAUNCHES[0]
    t0 = time.perf_counter()
    sim.evolve(2.0e-4, split=True, split_locs=[1])          # ~20 CFL-limited RK45 steps at dx = 1e-5
    torch.cuda.synchronize()
    t_ev = time.perf_counter() - t0
    launches = _lib_mod.LAUNCHES[0] - before
    nfev = sim._rk[1].nfev if sim._rk is not None else None
    attempts = (nfev - 2) // 6 if nfev else None
    out["rk45_evolve_sod_nx1e5"] = {"t_diff": 2.0e-4, "seconds": t_ev, "launches": launches, "rhs_evaluations": nfev,
                                    "attempted_steps": attempts,
                                    "ms_per_attempted_step": 1e3 * t_ev / attempts if attempts else None,
                                    "control": "scipy's RK45 step-size controller on the device, attempted steps replayed "
                                               "from a CUDA graph (17 launches each), flags read back per batch of <= 8",
                                    "finite": bool(np.isfinite(d.values()).all())}
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--members", type=int, default=65536, help="ensemble members per GPU")
    ap.add_argument("--e2e-steps", type=int, default=12)
    ap.add_argument("--cpu-sample", type=int, default=10240,
                    help="members the 1-core oracle-port baseline assembles (~10 s); the reference itself gets 1/160 of it")
    ap.add_argument("--cpu-newton-sample", type=int, default=8, help="steady states of the 1-core reference baseline")
    ap.add_argument("--ref-members", type=int, default=4096, help="members per step of the reference arm (oracle port)")
    ap.add_argument("--ref-members-per-core", type=int, default=1,
                    help="members per step and host core of the reference arm (unmodified reference, ~0.2 s each)")
    ap.add_argument("--newton-steps", type=int, default=5, help="timed steady-state passes of the Newton leg")
    ap.add_argument("--slab-nx", type=int, default=10_000_000)
    ap.add_argument("--slab-steps", type=int, default=200)
    ap.add_argument("--slab-newton-nx", type=int, default=1_000_000, help="cells per rank of the slab Newton leg")
    ap.add_argument("--no-slab", action="store_true")
    ap.add_argument("--no-extra", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
