"""AMR parity at BASELINE config 3's scale (~1e5 cells), against the UNMODIFIED reference's Refiner (refine.py:90-282).

A 1e5-cell mesh is too large to commit, and identical meshes are identical byte strings: the fixture stores, for three
consecutive refinement passes of the reference on a 100,000-cell Sod-like state (each pass inserts ~2e4 midpoints in the
three fronts and deletes 5e3-2e4 cells of the exactly flat far fields: 100,000 -> 98,683 -> 119,221 -> 129,044 cells),
the new size, SHA-256 digests of the positions and of the values, and every 997th cell.  The input is built from +, -, *,
min and max only (clamped cubic fronts, no libm calls), so tests/test_gpu_refine.py rebuilds it bit-identically on any
machine.

    python tests/golden/make_amr_large_golden.py     ->  tests/golden/amr_large.npz   (build container only)
"""
import contextlib
import hashlib
import io
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

NX0 = 100_000
CRITERIA = (2.0, 2e-5, 2e-5, 1e-6)         # ratio, slope, curve, prune (> 0: cells are also deleted)
PASSES = 3


def front(x, x0, k):
    """A smooth step from exactly 0 to exactly 1 over |x - x0| < 1/k (clamped cubic), in IEEE-exact operations only."""
    t = np.minimum(np.maximum(k * (x - x0), -1.0), 1.0)
    return 0.5 + (0.75 * t - 0.25 * (t * (t * t)))


def sod_like_state(nx=NX0):
    """Three wave fronts (rarefaction head, contact, shock) of different widths on a uniform mesh: [nx], [nx, 3]."""
    x = np.linspace(0.0, 1.0, nx)
    a, b, c = front(x, 0.35, 8.0), front(x, 0.62, 1000.0), front(x, 0.81, 10000.0)
    rho = 1.0 - 0.55 * a - 0.18 * b - 0.145 * c
    u = 0.9 * a - 0.9 * c
    p = 1.0 - 0.7 * a - 0.2 * c
    return x, np.stack([rho, rho * u, p / 0.4 + 0.5 * (rho * (u * u))], axis=1)


def digest(a):
    return hashlib.sha256(np.ascontiguousarray(a, dtype=np.float64).tobytes()).hexdigest()


def main():
    from baseline import reference_env
    reference_env.activate()
    from splitfxm.domain import Domain
    from splitfxm.refine import Refiner
    x, U = sod_like_state()
    d = Domain.from_size(NX0, 1, 1, ["rho", "rhou", "E"])
    for cell, row in zip(d.interior(), U):
        cell.set_values(np.array(row))
    r = Refiner()
    r.set_criteria(*CRITERIA)
    r.set_max_points(10 * NX0)
    out = {"nx0": NX0, "criteria": np.array(CRITERIA), "input_x_sha": digest(x), "input_U_sha": digest(U)}
    for k in range(PASSES):
        with contextlib.redirect_stdout(io.StringIO()):
            r.refine(d)
        xs = np.array(d.positions(interior=True))
        Us = np.array(d.values(interior=True))
        out[f"n{k}"], out[f"x_sha{k}"], out[f"U_sha{k}"] = xs.shape[0], digest(xs), digest(Us)
        out[f"x_sample{k}"], out[f"U_sample{k}"] = xs[::997].copy(), Us[::997].copy()
        print("pass", k, "cells", xs.shape[0])
    np.savez_compressed(os.path.join(HERE, "amr_large.npz"), **out)


if __name__ == "__main__":
    main()
