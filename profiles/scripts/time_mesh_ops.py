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
