"""One invocation of each secondary hot-path operation after a warm-up, for `ncu --metrics gpu__time_duration.sum`
launch lists (profiles/r2_*_launches.csv).  Usage: python profiles/scripts/prof_ops.py <op> [<op> ...]
ops: newton_ensemble | newton_mesh | amr | residual_mesh | jacobian_mesh | euler_sod | rk45_sod"""
import contextlib
import io
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import splitfxm_b200 as sf  # noqa: E402
from splitfxm_b200.ensemble import EnsembleSimulation  # noqa: E402
from splitfxm_b200.models.advection_diffusion import AdvectionDiffusion  # noqa: E402
from splitfxm_b200.models.euler1d import Euler1DConservative  # noqa: E402

ops = sys.argv[1:] or ["newton_ensemble"]
dev = torch.device("cuda:0")
torch.cuda.set_device(0)
marker = torch.zeros(1, device=dev)


def mark():
    """A recognisable launch between the warm-up and the measured invocation (torch fill kernel on 1 element)."""
    marker.fill_(1.0)
    torch.cuda.synchronize()


if "newton_ensemble" in ops:
    B = 65536
    x, y0 = bench.initial_state()
    c, nu = bench.member_params(0, B, B)
    d = sf.Domain.from_size(bench.NX, 1, 1, bench.COMPS)
    ens = EnsembleSimulation(d, "ad", {"c": c, "nu": nu}, bench.BCS, sf.FDSchemes.CENTRAL, method="FDM", device=dev)
    Y = torch.from_numpy(bench.member_states(y0, 0, B)).to(dev)
    ens.steady_state(Y, split=True, split_locs=[1])
    torch.cuda.synchronize()
    mark()
    ens.steady_state(Y, split=True, split_locs=[1])
    torch.cuda.synchronize()
    del ens, Y

if "newton_mesh" in ops or "residual_mesh" in ops or "jacobian_mesh" in ops:
    nx = 1_000_000
    d = sf.Domain.from_size(nx, 1, 1, bench.COMPS)
    sim = sf.Simulation(d, AdvectionDiffusion(0.2, 0.001, "FDM"), bench.ICS, bench.BCS, sf.FDSchemes.CENTRAL)
    e1 = sim._engine()
    y = sim._state(e1, True, [1])
    Jb = e1.jacobian_buffer()
    rr = torch.empty_like(y)
    ss = torch.zeros_like(y)
    st1 = torch.zeros(1, dtype=torch.int32, device=dev)

    def it():
        e1.jacobian(y, True, [1], J=Jb, r0=rr)
        e1.solve(Jb, rr, ss, True, [1], status=st1)
        e1.update_norms(y, ss)
    it()
    e1.residual(y, rr, True, [1])
    torch.cuda.synchronize()
    mark()
    if "residual_mesh" in ops:
        e1.residual(y, rr, True, [1])
    if "jacobian_mesh" in ops:
        e1.jacobian(y, True, [1], J=Jb)
    if "newton_mesh" in ops:
        it()
    torch.cuda.synchronize()

if "amr" in ops or "euler_sod" in ops or "rk45_sod" in ops:
    comps = ["rho", "rhou", "E"]
    nbc = {c: {"left": {"neumann": 0.0}, "right": {"neumann": 0.0}} for c in comps}
    nx = int(os.environ.get("PROF_NX", "1000000"))
    d = sf.Domain.from_size(nx, 1, 1, comps)
    xs = d.interior_x()
    U = d.interior_U()
    U[:, 0] = np.where(xs < 0.5, 1.0, 0.125) + 0.01 * np.tanh(200 * (xs - 0.3))
    U[:, 2] = np.where(xs < 0.5, 2.5, 0.25)
    sim = sf.Simulation(d, Euler1DConservative(1.4, "FVM"), {}, nbc, sf.FVSchemes.LAX_FRIEDRICHS)
    sim._r.set_criteria(2.0, 0.05, 0.05, -0.001)
    sim._r.set_max_points(4 * nx)
    if "euler_sod" in ops:
        dt = 0.4 * (xs[1] - xs[0]) / 2.2
        sim.evolve(3 * dt, method="Euler", max_step=dt)
        torch.cuda.synchronize()
        mark()
        sim.evolve(3 * dt, method="Euler", max_step=dt)
        torch.cuda.synchronize()
    if "rk45_sod" in ops:
        sim.evolve(2e-6)
        torch.cuda.synchronize()
        mark()
        sim.evolve(2e-6)
        torch.cuda.synchronize()
    if "amr" in ops:
        with contextlib.redirect_stdout(io.StringIO()):
            d2 = sf.Domain.from_arrays(d._host()[0], d._host()[1], 1, 1, comps)
            sim._r.refine(d2)
            torch.cuda.synchronize()
            mark()
            sim._r.refine(d)
        torch.cuda.synchronize()
print("done", ops)
