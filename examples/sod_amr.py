#!/usr/bin/env python
"""BASELINE config 3: conservative Sod shock tube (Euler1DConservative, FVM Lax-Friedrichs, neumann-0 ends) with
adaptive mesh refinement, mesh grown to ~1e5 cells.

Recipe (from a 100-cell base the reference's criteria saturate near ~220 cells because only the three
wave fronts refine, so the run starts from a large base mesh and uses criteria tight enough to also split the cells
of the rarefaction fan):
    base mesh 60,000 cells; evolve(0.02) with RK45 (the reference's default) so the three waves have separated;
    set_criteria(ratio=2.0, slope=2e-5, curve=2e-5, prune=-0.001), set_max_points(400,000);
    repeat: Refiner.refine -> evolve(1e-5), until >= 1e5 cells or 8 passes.
Every call goes through the reference-shaped API (Simulation.evolve, s._r.refine(s._d)); prints one JSON line.
"""
import contextlib
import io
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def main():
    import splitfxm_b200 as sf
    from splitfxm_b200 import _lib
    from splitfxm_b200.models.euler1d import Euler1DConservative
    names = ["rho", "rhou", "E"]
    bcs = {n: {"left": {"neumann": 0.0}, "right": {"neumann": 0.0}} for n in names}
    d = sf.Domain.from_size(60_000, 1, 1, names)
    m = Euler1DConservative(gamma=1.4, method="FVM")
    s = sf.Simulation(d, m, {}, bcs, sf.default_scheme("FVM"))
    xs = d.interior_x()
    d.interior_U()[:, 0] = np.where(xs < 0.5, 1.0, 0.125)
    d.interior_U()[:, 2] = np.where(xs < 0.5, 2.5, 0.25)
    s._r.set_criteria(2.0, 2e-5, 2e-5, -0.001)
    s._r.set_max_points(400_000)
    history = [d._nx]
    t_evolve = t_refine = 0.0
    t_sim = 0.0
    passes = 0
    _lib.LAUNCHES[0] = 0
    t0 = time.perf_counter()
    s.evolve(0.02, split=True, split_locs=[1])
    t_evolve += time.perf_counter() - t0
    t_sim += 0.02
    while d._nx < 100_000 and passes < 8:
        t0 = time.perf_counter()
        with contextlib.redirect_stdout(io.StringIO()):
            s._r.refine(s._d)
        t_refine += time.perf_counter() - t0
        history.append(d._nx)
        passes += 1
        t0 = time.perf_counter()
        s.evolve(1e-5, split=True, split_locs=[1])
        t_evolve += time.perf_counter() - t0
        t_sim += 1e-5
    U = d.values(interior=True)
    x = np.array(d.positions(interior=True))
    print(json.dumps({"config": "Sod shock tube, Euler1DConservative FVM LF, RK45 + AMR", "cells": history,
                      "passes": passes, "t_sim": t_sim, "evolve_seconds": t_evolve, "refine_seconds": t_refine,
                      "gpu_launches": _lib.LAUNCHES[0], "min_dx": float(np.min(np.diff(x))),
                      "mesh_sorted": bool(np.all(np.diff(x) > 0)), "finite": bool(np.isfinite(U).all()),
                      "rho_min_max": [float(U[:, 0].min()), float(U[:, 0].max())]}))


if __name__ == "__main__":
    main()
