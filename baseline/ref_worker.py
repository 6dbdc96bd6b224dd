"""Worker functions that run the UNMODIFIED reference (baseline/_ref) on the bench workload.

Each worker process (multiprocessing 'spawn': no CUDA state, no product modules) activates the reference
environment once, builds the reference `Simulation`s of its members exactly as benchmark/test_jac_benchmark.py:14-56
does (Domain.from_size(250, 1, 1, ...), AdvectionDiffusion(c, nu, 'FDM'), FDSchemes.CENTRAL) and then times

  assemble : s.jacobian(x0, True, [1])                    (simulation.py:270-394)
  steady   : s.steady_state(split=True, split_locs=[1])   (simulation.py:474-526; the solver is the oracle/newton.py
             stand-in for the absent third-party `splitnewton` - parity unpinned)

Used only by bench.py's `--impl reference` arm and its `cpu_baseline` leg.
"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

_state = {}


def _activate():
    if "mods" not in _state:
        os.environ.setdefault("OMP_NUM_THREADS", "1")
        os.environ.setdefault("OPENBLAS_NUM_THREADS", "1")
        from baseline import reference_env
        reference_env.activate()
        from splitfxm.domain import Domain
        from splitfxm.simulation import Simulation
        from splitfxm.schemes import FDSchemes
        from splitfxm.models.advection_diffusion import AdvectionDiffusion
        _state["mods"] = (Domain, Simulation, FDSchemes, AdvectionDiffusion)
    return _state["mods"]


def _members(first, count, total):
    """Reference Simulations + initial vectors of members [first, first + count) of the bench ensemble."""
    key = (first, count, total)
    if _state.get("key") != key:
        import bench
        Domain, Simulation, FDSchemes, AdvectionDiffusion = _activate()
        c, nu = bench.member_params(first, count, total)
        sims, y0 = [], None
        for k in range(count):
            d = Domain.from_size(bench.NX, 1, 1, bench.COMPS)
            s = Simulation(d, AdvectionDiffusion(float(c[k]), float(nu[k]), "FDM"), bench.ICS, bench.BCS, FDSchemes.CENTRAL)
            if y0 is None:
                y0 = d.listify_interior(bench.SPLIT, bench.LOCS)
            sims.append(s)
        Y = bench.member_states(y0, first, count)
        _state["key"], _state["sims"], _state["Y"] = key, sims, Y
    return _state["sims"], _state["Y"]


def prepare(args):
    _members(*args)
    return True


def assemble(args):
    """-> (seconds, checksum): one Jacobian assembly of every member of the slice."""
    import bench
    sims, Y = _members(*args)
    acc = 0.0
    t0 = time.perf_counter()
    for k, s in enumerate(sims):
        J = s.jacobian(Y[:, k], bench.SPLIT, bench.LOCS)
        acc += float(J[0, 0])
    return time.perf_counter() - t0, acc


def steady(args):
    """-> (seconds, iterations): one Split-Newton steady state of every member of the slice."""
    import bench
    sims, Y = _members(*args)
    its = []
    t0 = time.perf_counter()
    for k, s in enumerate(sims):
        s.initialize_from_list(Y[:, k], bench.SPLIT, bench.LOCS)
        its.append(int(s.steady_state(split=bench.SPLIT, split_locs=bench.LOCS)))
    return time.perf_counter() - t0, its


# ---------------------------------------------------------------------------------------------------------
# a fixed set of worker processes, one slice of members each (bench.py's reference arm)
# ---------------------------------------------------------------------------------------------------------
def _serve(conn, slice_args):
    try:
        prepare(slice_args)
        conn.send(("ready", None))
        while True:
            cmd = conn.recv()
            if cmd == "stop":
                break
            conn.send((cmd, assemble(slice_args) if cmd == "assemble" else steady(slice_args)))
    except Exception as exc:                                   # report instead of hanging the parent
        conn.send(("error", repr(exc)))


class RefPool:
    """`procs` processes; process i owns members [first + i * per_proc, first + (i + 1) * per_proc)."""

    def __init__(self, procs, per_proc, total, first=0):
        import multiprocessing as mp
        ctx = mp.get_context("spawn")
        self.workers = []
        for i in range(procs):
            parent, child = ctx.Pipe()
            p = ctx.Process(target=_serve, args=(child, (first + i * per_proc, per_proc, total)), daemon=True)
            p.start()
            self.workers.append((p, parent))
        for _, conn in self.workers:
            tag, msg = conn.recv()
            if tag != "ready":
                raise RuntimeError("reference worker failed: " + str(msg))

    def step(self, cmd):
        """Run `cmd` ('assemble' | 'steady') on every worker; -> (wall seconds, [results])."""
        t0 = time.perf_counter()
        for _, conn in self.workers:
            conn.send(cmd)
        out = []
        for _, conn in self.workers:
            tag, msg = conn.recv()
            if tag == "error":
                raise RuntimeError("reference worker failed: " + str(msg))
            out.append(msg)
        return time.perf_counter() - t0, out

    def close(self):
        for p, conn in self.workers:
            try:
                conn.send("stop")
            except Exception:
                pass
        for p, _ in self.workers:
            p.join(timeout=10)
