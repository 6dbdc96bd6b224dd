"""Diagnostic (not a test): how the CUDA Euler Jacobian's zero pattern differs from the reference fixtures."""
import json
import sys
import os
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, ROOT)
from helpers import RESIDUAL_CASES, load, product_simulation

out = {}
for name in RESIDUAL_CASES:
    if not name.startswith("sod_"):
        continue
    z = load(name)
    s, d, m = product_simulation(z)
    J = s.jacobian(z["x0"], m["split"], m["locs"]).tocoo()
    n = J.shape[0]
    ref = np.zeros((n, n)); ref[z["jac_row"], z["jac_col"]] = z["jac_val"]
    got = np.zeros((n, n)); got[J.row, J.col] = J.data
    only_ref = (ref != 0) & (got == 0)
    only_got = (ref == 0) & (got != 0)
    scale = np.max(np.abs(ref))
    r = s.get_residuals_from_list(z["x0"], m["split"], m["locs"])
    rmax = np.max(np.abs(z["residual"]))
    out[name] = dict(n=n, nnz_ref=int((ref != 0).sum()), nnz_got=int((got != 0).sum()),
                     only_ref=int(only_ref.sum()), only_got=int(only_got.sum()),
                     max_abs_only_ref=float(np.abs(ref[only_ref]).max()) if only_ref.any() else 0.0,
                     max_abs_only_got=float(np.abs(got[only_got]).max()) if only_got.any() else 0.0,
                     jac_scale=float(scale), max_abs_diff=float(np.abs(ref - got).max()),
                     res_max=float(rmax), res_max_abs_diff=float(np.abs(r - z["residual"]).max()),
                     res_bit_equal=bool(np.array_equal(r, z["residual"])),
                     smallest_common=float(np.abs(ref[(ref != 0) & (got != 0)]).min()))
print(json.dumps(out, indent=1))
