"""Shared helpers: build oracle Problems and splitfxm_b200 Simulations from the golden fixtures."""
import glob
import json
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

RESIDUAL_CASES = sorted(os.path.basename(p)[:-4] for p in
                        glob.glob(os.path.join(GOLDEN, "ad_*.npz")) + glob.glob(os.path.join(GOLDEN, "sod_*.npz")))
EVOLVE_CASES = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN, "evolve_*.npz")))


def load(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"))


def case_meta(z):
    model = str(z["model"])
    kw = dict(c=float(z["c"]), nu=float(z["nu"])) if model == "ad" else dict(gamma=float(z["gamma"]))
    split = bool(z["split"])
    locs = [int(v) for v in z["locs"]] if split else None
    return dict(model=model, kw=kw, method=str(z["method"]), scheme=str(z["scheme"]), nx=int(z["nx"]),
                nbl=int(z["nbl"]), nbr=int(z["nbr"]), comps=[str(c) for c in z["comps"]],
                bcs=json.loads(str(z["bcs"])), split=split, locs=locs, xmin=float(z["xmin"]), xmax=float(z["xmax"]))


def oracle_problem(z, x_all=None):
    from oracle import core
    m = case_meta(z)
    x_all = z["x_all"] if x_all is None else x_all
    x_int = x_all[m["nbl"]:m["nbl"] + m["nx"]]
    return core.Problem(core.make_model(m["model"], **m["kw"]), m["method"], m["scheme"], x_int, m["nbl"], m["nbr"],
                        m["bcs"], m["comps"], m["xmin"], m["xmax"]), m


def product_simulation(z, x_all=None, U_all=None):
    """splitfxm_b200 Simulation whose Domain holds the fixture's state (ghosts included)."""
    import splitfxm_b200 as sf
    from splitfxm_b200.models.advection_diffusion import AdvectionDiffusion
    from splitfxm_b200.models.euler1d import Euler1D, Euler1DConservative
    m = case_meta(z)
    x_all = z["x_all"] if x_all is None else x_all
    U_all = z["U_all"] if U_all is None else U_all
    if m["model"] == "ad":
        model = AdvectionDiffusion(m["kw"]["c"], m["kw"]["nu"], m["method"])
    elif m["model"] == "euler_prim":
        model = Euler1D(m["kw"]["gamma"], m["method"])
    else:
        model = Euler1DConservative(m["kw"]["gamma"], m["method"])
    d = sf.Domain.from_arrays(x_all, U_all, m["nbl"], m["nbr"], m["comps"], m["xmin"], m["xmax"])
    scheme = (sf.FVSchemes if m["method"] == "FVM" else sf.FDSchemes)[m["scheme"]]
    s = sf.Simulation(d, model, {}, m["bcs"], scheme)
    return s, d, m


def rel_err(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    den = np.max(np.abs(b))
    return float(np.max(np.abs(a - b)) / (den if den > 0 else 1.0))
