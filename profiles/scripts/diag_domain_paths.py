"""Diagnostic: where do the device-resident and the host-synced sequences diverge?"""
import contextlib, io, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import splitfxm_b200 as sf
from splitfxm_b200.models.advection_diffusion import AdvectionDiffusion
comps = ["u", "v", "w"]
bcs = {"u": {"left": {"dirichlet": 1.0}, "right": {"dirichlet": 0.0}}, "v": {"left": {"dirichlet": 3}, "right": {"dirichlet": 4}},
       "w": {"left": {"dirichlet": 2}, "right": {"neumann": 0.0}}}
def make():
    d = sf.Domain.from_size(200, 1, 1, comps)
    s = sf.Simulation(d, AdvectionDiffusion(0.2, 0.05, "FVM"), {"u": "gaussian", "v": "rarefaction", "w": "tophat"}, bcs, sf.FVSchemes.LAX_FRIEDRICHS)
    s._r.set_criteria(3.0, 0.05, 0.05, -0.001)
    return d, s
def snap(d):
    dx, dU = d._device()
    return dx.cpu().numpy().copy(), dU.cpu().numpy().copy()
d1, s1 = make(); d2, s2 = make()
def cmp(tag):
    a, b = snap(d1), snap(d2)
    same_x = a[0].shape == b[0].shape and np.array_equal(a[0], b[0]); same_U = a[1].shape == b[1].shape and np.array_equal(a[1], b[1])
    print(tag, "nx", d1._nx, d2._nx, "x equal", same_x, "U equal", same_U)
    if a[1].shape == b[1].shape and not same_U:
        w = np.argwhere(a[1] != b[1]); print("   first diffs", w[:5].tolist(), a[1][tuple(w[0])], b[1][tuple(w[0])], "count", len(w))
cmp("init")
s1.evolve(0.01); s2.evolve(0.01); d2.interior_U(); cmp("after rk45")
with contextlib.redirect_stdout(io.StringIO()):
    s1._r.refine(d1); s2._r.refine(d2)
d2.interior_x(); cmp("after refine")
s1.evolve(0.01, method="Euler", max_step=1e-4); s2.evolve(0.01, method="Euler", max_step=1e-4); d2.values(); cmp("after euler")
i1 = s1.steady_state(split=True, split_locs=[1, 2]); i2 = s2.steady_state(split=True, split_locs=[1, 2]); cmp("after steady %d %d" % (i1, i2))
