"""Diagnostic (not a test): is the custom Burgers model bit-identical to the oracle's NumPy evaluation?"""
import json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, ROOT)
from nonlinear_cases import CASES, oracle_case, oracle_solve, product_case
from oracle import core
out = {}
for name in ("burgers_fvm_lf_plain", "burgers_fdm_central_armijo", "euler_duct_plain"):
    P, U0, c = oracle_case(name)
    s, d, _ = product_case(name)
    x0 = core.to_flat(U0, c["split"], c["locs"])
    r = s.get_residuals_from_list(x0, c["split"], c["locs"]); rr = P.residual(x0, c["split"], c["locs"])
    J = s.jacobian(x0, c["split"], c["locs"]).toarray(); Jr = P.jacobian_matrix(x0, c["split"], c["locs"]).toarray()
    out[name] = dict(res_bit_equal=bool(np.array_equal(r, rr)), res_max_diff=float(np.abs(r - rr).max()), res_max=float(np.abs(rr).max()),
                     jac_bit_equal=bool(np.array_equal(J, Jr)), jac_max_diff=float(np.abs(J - Jr).max()), jac_max=float(np.abs(Jr).max()),
                     pattern_equal=bool(np.array_equal(J != 0, Jr != 0)))
print(json.dumps(out, indent=1))
