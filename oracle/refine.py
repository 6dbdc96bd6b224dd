"""Restatement of the reference's adaptive mesh refinement, Refiner.refine / perform_changes (refine.py:90-282).

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Works on arrays: ``x[n]`` interior positions, ``U[n, nv]`` interior
values.  The marking rules are restated loop for loop with the reference's dictionaries (``keep``: 1 stays / -1 goes,
``loc``: insert a midpoint after that cell), so that the order-dependent parts - ``elif j not in keep`` (refine.py:169,
195), the adjacent-deletion loop (refine.py:231-233) and the delete-while-enumerating loop of perform_changes
(refine.py:273-275) - behave exactly as in the reference.  Pinned against the meshes the unmodified reference produced
(tests/golden/refine_random.npz, refine_sod.npz, amr_large.npz) in tests/test_oracle_golden.py."""
from __future__ import annotations

import math

import numpy as np

EPS = math.ulp(1.0)          # refine.py:7


class RefineError(Exception):
    """Mirrors splitfxm.error.SFXM raised by the Refiner (refine.py:57-70, 123)."""


class Refiner:
    def __init__(self):                                      # refine.py:20-36
        self.ratio, self.slope, self.curve, self.prune = 10.0, 0.8, 0.8, -0.001
        self.npmax = 1000
        self.min_range = 0.01
        self.min_grid = 1e-10

    def set_criteria(self, ratio, slope, curve, prune):      # refine.py:38-75
        if ratio < 2.0:
            raise RefineError(f"ratio must be greater than 2.0 ({ratio} was specified).")
        elif slope < 0.0 or slope > 1.0:
            raise RefineError(f"slope must be between 0.0 and 1.0 ({slope} was specified).")
        elif curve < 0.0 or curve > 1.0:
            raise RefineError(f"curve must be between 0.0 and 1.0 ({curve} was specified).")
        elif prune > curve or prune > slope:
            raise RefineError(f"prune must be less than 'curve' and 'slope' ({prune} was specified).")
        self.ratio, self.slope, self.curve, self.prune = ratio, slope, curve, prune

    def set_max_points(self, npmax):                         # refine.py:77-87
        self.npmax = npmax

    # refine.py:90-236: the marking
    def mark(self, x, U):
        """-> (loc, keep) dictionaries exactly as Refiner.refine builds them."""
        x = np.asarray(x, dtype=np.float64)
        U = np.asarray(U, dtype=np.float64)
        n, nv = U.shape
        keep, loc = {}, {}
        keep[0] = 1                                          # refine.py:119-120
        keep[n - 1] = 1
        if n > self.npmax:                                   # refine.py:122-123
            raise RefineError("Exceeded maximum number of points")
        z = [float(v) for v in x]
        dz = [z[i + 1] - z[i] for i in range(n - 1)]         # refine.py:126
        for i in range(nv):
            v = [float(t) for t in U[:, i]]
            s = [(v[j + 1] - v[j]) / (z[j + 1] - z[j]) for j in range(n - 1)]     # refine.py:138-142
            vmin, vmax = min(v), max(v)
            smin, smax = min(s), max(s)
            aa = max(abs(vmin), abs(vmax))
            ss = max(abs(smin), abs(smax))
            if (vmax - vmin) > self.min_range * aa:          # refine.py:158-170
                dmax = self.slope * (vmax - vmin) + EPS
                for j in range(n - 1):
                    r = abs(v[j + 1] - v[j]) / dmax
                    if r > 1.0 and dz[j] >= 2 * self.min_grid:
                        loc[j] = 1
                    if r >= self.prune:
                        keep[j] = 1
                        keep[j + 1] = 1
                    elif j not in keep:
                        keep[j] = -1
            if (smax - smin) > self.min_range * ss:          # refine.py:176-196
                dmax = self.curve * (smax - smin)
                for j in range(n - 2):
                    r = abs(s[j + 1] - s[j]) / (dmax + (EPS / dz[j]))
                    if r > 1.0 and dz[j] >= 2 * self.min_grid and dz[j + 1] >= 2 * self.min_grid:
                        loc[j] = 1
                        loc[j + 1] = 1
                    if r >= self.prune:
                        keep[j + 1] = 1
                    elif (j + 1) not in keep:
                        keep[j + 1] = -1
        for j in range(1, n - 1):                            # refine.py:199-227: the grid's own ratios
            if dz[j] > self.ratio * dz[j - 1]:
                loc[j] = 1
                keep[j - 1] = 1
                keep[j] = 1
                keep[j + 1] = 1
                keep[j + 2] = 1
            if dz[j] < dz[j - 1] / self.ratio:
                loc[j - 1] = 1
                keep[j - 2] = 1
                keep[j - 1] = 1
                keep[j] = 1
                keep[j + 1] = 1
            if j > 1 and z[j + 1] - z[j - 1] > self.ratio * dz[j - 2]:
                keep[j] = 1
            if j < n - 2 and z[j + 1] - z[j - 1] > self.ratio * dz[j + 1]:
                keep[j] = 1
        for j in range(2, n - 1):                            # refine.py:231-233: never two adjacent deletions
            if j in keep and j - 1 in keep and keep[j] == -1 and keep[j - 1] == -1:
                keep[j] = 1
        return loc, keep

    # refine.py:239-282
    @staticmethod
    def perform_changes(x, U, loc, keep):
        """-> (x_new, U_new).  Deletion by `del cells[i]` while enumerating (refine.py:273-275) skips the cell that
        follows a deleted one; new cells are plain averages (refine.py:268-269); the merged list is sorted by x
        (cell.py:28-58: cells compare on x; Python's sort is stable, so equal positions keep list order)."""
        x = np.asarray(x, dtype=np.float64)
        U = np.asarray(U, dtype=np.float64)
        n = x.shape[0]
        marked = [bool(i in keep and keep[i] == -1) for i in range(n)]
        new_x = [0.5 * (x[i + 1] + x[i]) for i in loc.keys()]
        new_U = [0.5 * (U[i + 1] + U[i]) for i in loc.keys()]
        cells = list(range(n))                               # indices into the old arrays
        i = 0
        while i < len(cells):                                # enumerate() over a list that shrinks under it
            if marked[cells[i]]:
                del cells[i]
            i += 1
        items = [(float(x[k]), U[k]) for k in cells] + list(zip([float(v) for v in new_x], new_U))
        items.sort(key=lambda t: t[0])
        xs = np.array([t[0] for t in items], dtype=np.float64)
        Us = np.array([t[1] for t in items], dtype=np.float64).reshape(len(items), U.shape[1])
        return xs, Us

    def refine(self, x, U):
        loc, keep = self.mark(x, U)
        return self.perform_changes(x, U, loc, keep)
