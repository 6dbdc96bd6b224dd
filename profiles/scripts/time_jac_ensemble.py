"""CUDA-event times of the ensemble Jacobian, residual and fused Newton linear step on the 65,536-member bench
ensemble - for A/B runs of library builds (SFXM_LIB=<variant>)."""
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
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


J = ens.jacobian(Y, bench.SPLIT, bench.LOCS)
out = torch.empty_like(Y)
msj = min(timed(lambda: eng.jacobian(Y, bench.SPLIT, bench.LOCS, J=J), 20) for _ in range(3))
msr = min(timed(lambda: eng.residual(Y, out, bench.SPLIT, bench.LOCS), 20) for _ in range(3))
s = torch.zeros_like(Y)
status = torch.zeros(B, dtype=torch.int32, device=dev)
ms0 = timed(lambda: eng.newton_solve_fused(Y, s, True, [1], 0, status=status), 5)
ms1 = timed(lambda: eng.newton_solve_fused(Y, s, True, [1], 1, status=status), 5)
print(os.environ.get("SFXM_LIB", "default"), "jacobian %.4f ms  residual %.4f ms  fused solve seg0 %.4f seg1 %.4f ms  checksum %.17g"
      % (msj, msr, ms0, ms1, float(J.sum())))
