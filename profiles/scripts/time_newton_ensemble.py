"""CUDA-event time of one Split-Newton steady state of the 65,536-member bench ensemble, and of the two fused linear
steps alone - for A/B runs of library builds (SFXM_LIB=<variant>)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import splitfxm_b200 as sf  # noqa: E402
from splitfxm_b200.ensemble import EnsembleSimulation  # noqa: E402

dev = torch.device("cuda:0")
B = 65536
x, y0 = bench.initial_state()
c, nu = bench.member_params(0, B, B)
d = sf.Domain.from_size(bench.NX, 1, 1, bench.COMPS)
ens = EnsembleSimulation(d, "ad", {"c": c, "nu": nu}, bench.BCS, sf.FDSchemes.CENTRAL, method="FDM", device=dev)
Y = torch.from_numpy(bench.member_states(y0, 0, B)).to(dev)
eng = ens.eng


def timed(fn, reps):
    for _ in range(3):
        r = fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        r = fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


s = torch.zeros_like(Y)
status = torch.zeros(B, dtype=torch.int32, device=dev)
ms0 = timed(lambda: eng.newton_solve_fused(Y, s, True, [1], 0, status=status), 5)
ms1 = timed(lambda: eng.newton_solve_fused(Y, s, True, [1], 1, status=status), 5)
ms = timed(lambda: ens.steady_state(Y, split=True, split_locs=[1]), 8)
print(os.environ.get("SFXM_LIB", "default"), "fused solve seg0 %.4f ms seg1 %.4f ms | steady state %.3f ms" % (ms0, ms1, ms))
