"""CUDA-event times of the one-mesh residual / Jacobian / Newton iteration at a given nx (default 1e6), for A/B runs of
library builds (SFXM_LIB=<path to a variant libsfxm_b200.so>)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import splitfxm_b200 as sf  # noqa: E402
from splitfxm_b200.models.advection_diffusion import AdvectionDiffusion  # noqa: E402

nx = int(os.environ.get("PROF_NX", "1000000"))
d = sf.Domain.from_size(nx, 1, 1, bench.COMPS)
sim = sf.Simulation(d, AdvectionDiffusion(0.2, 0.001, "FDM"), bench.ICS, bench.BCS, sf.FDSchemes.CENTRAL)
e1 = sim._engine()
y = sim._state(e1, True, [1])
Jb = e1.jacobian_buffer()
rr = torch.empty_like(y)


def timed(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


ss = torch.zeros_like(y)
st1 = torch.zeros(1, dtype=torch.int32, device=y.device)


def newton_iteration(split, locs):
    def it():
        e1.jacobian(y, split, locs, J=Jb, r0=rr)
        e1.solve(Jb, rr, ss, split, locs, status=st1)
        e1.update_norms(y, ss, update=False)
    return it


print(os.environ.get("SFXM_LIB", "default"), "nx", nx, "jacobian_ms %.4f residual_ms %.4f newton_it_split1_ms %.4f newton_it_unsplit_ms %.4f" % (
    timed(lambda: e1.jacobian(y, True, [1], J=Jb)), timed(lambda: e1.residual(y, rr, True, [1])),
    timed(newton_iteration(True, [1]), 10), timed(newton_iteration(False, None), 10)))

# conservative Sod, explicit Euler step (residual_warp_kernel with the fused update) and one AMR pass, nx = 1e7
import numpy as np  # noqa: E402
from splitfxm_b200 import _lib as _lib_mod  # noqa: E402
from splitfxm_b200.models.euler1d import Euler1DConservative  # noqa: E402

nx2 = int(os.environ.get("PROF_NX2", "10000000"))
comps = ["rho", "rhou", "E"]
nbc = {c: {"left": {"neumann": 0.0}, "right": {"neumann": 0.0}} for c in comps}
d2 = sf.Domain.from_size(nx2, 1, 1, comps)
xs = d2.interior_x()
U = d2.interior_U()
U[:, 0] = np.where(xs < 0.5, 1.0, 0.125) + 0.01 * np.tanh(200 * (xs - 0.3))
U[:, 2] = np.where(xs < 0.5, 2.5, 0.25)
sim2 = sf.Simulation(d2, Euler1DConservative(1.4, "FVM"), {}, nbc, sf.FVSchemes.LAX_FRIEDRICHS)
e2 = sim2._engine()
y2 = sim2._state(e2, False, None)
o2 = torch.empty_like(y2)
dt = 0.4 * float(xs[1] - xs[0]) / 2.2
ms_euler = timed(lambda: e2.euler_step(y2, o2, dt), 20)
lib = _lib_mod.load()
dx, dU = d2.device_arrays()
xd, Ud = dx[1:-1].contiguous(), dU[1:-1].contiguous()
work = torch.empty(lib.sfxm_refine_fused_workspace_bytes(nx2), dtype=torch.uint8, device=xd.device)
counts = torch.zeros(3, dtype=torch.int64, device=xd.device)
xo = torch.empty(2 * nx2, dtype=torch.float64, device=xd.device)
Uo = torch.empty((2 * nx2, 3), dtype=torch.float64, device=xd.device)


def amr_pass():
    _lib_mod.check(lib.sfxm_refine_fused(xd.data_ptr(), Ud.data_ptr(), nx2, 3, 2.0, 0.05, 0.05, -0.001, 0.01, 1e-10,
                                         work.data_ptr(), xo.data_ptr(), Uo.data_ptr(), counts.data_ptr(), _lib_mod.stream_ptr()))


print("   nx2", nx2, "euler_step_ms %.4f (%.2f TB/s)  amr_pass_ms %.4f" % (ms_euler, nx2 * 56 / ms_euler / 1e9, timed(amr_pass, 10)))
