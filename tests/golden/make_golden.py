"""Generate golden fixtures from the UNMODIFIED reference (gpavanb1/SplitFXM at /root/reference).

Run in the build container only (the GPU box has no /root/reference):

    python tests/golden/make_golden.py

What it does
  1. copies /root/reference to a scratch dir and builds its two Cython modules there
     (`python setup.py build_ext --inplace`); nothing is written into /root/reference or this repo
     except the fixture files below;
  2. puts throw-away stubs for the three modules the reference imports but this image lacks
     (numdifftools, matplotlib, splitnewton) on sys.path - none of them is on a path exercised here;
  3. runs the reference's own public API on the BASELINE.json shapes and writes the inputs/outputs to
     tests/golden/*.npz.

Every fixture records the reference call that produced it (file:line) in its `what` field.
"""
import contextlib
import io
import os
import shutil
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"


def build_reference():
    work = os.path.join(tempfile.gettempdir(), "sfxm_ref_golden")
    if not os.path.exists(os.path.join(work, "splitfxm", "simulation.py")):
        shutil.copytree(REF, work, dirs_exist_ok=True)
    import glob
    if not glob.glob(os.path.join(work, "splitfxm", "flux*.so")):
        subprocess.check_call([sys.executable, "setup.py", "build_ext", "--inplace"], cwd=work,
                              stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    stubs = os.path.join(work, "_stubs")
    os.makedirs(os.path.join(stubs, "numdifftools"), exist_ok=True)
    os.makedirs(os.path.join(stubs, "matplotlib"), exist_ok=True)
    os.makedirs(os.path.join(stubs, "splitnewton"), exist_ok=True)
    with open(os.path.join(stubs, "numdifftools", "__init__.py"), "w") as f:
        f.write("import numpy as np\n"
                "class Jacobian:\n"
                "    def __init__(self, f, method='forward', step=1e-8):\n"
                "        self.f, self.h = f, step\n"
                "    def __call__(self, x):\n"
                "        x = np.asarray(x, dtype=float); f0 = np.asarray(self.f(x))\n"
                "        J = np.zeros((f0.size, x.size))\n"
                "        for k in range(x.size):\n"
                "            xp = x.copy(); xp[k] += self.h\n"
                "            J[:, k] = (np.asarray(self.f(xp)) - f0) / self.h\n"
                "        return J\n")
    with open(os.path.join(stubs, "matplotlib", "__init__.py"), "w") as f:
        f.write("")
    with open(os.path.join(stubs, "matplotlib", "pyplot.py"), "w") as f:
        f.write("def plot(*a, **k):\n    pass\n")
    with open(os.path.join(stubs, "splitnewton", "__init__.py"), "w") as f:
        f.write("")
    for mod, fn in (("newton", "newton"), ("split_newton", "split_newton")):
        with open(os.path.join(stubs, "splitnewton", mod + ".py"), "w") as f:
            f.write(f"def {fn}(*a, **k):\n    raise RuntimeError('splitnewton is not available')\n")
    sys.path.insert(0, stubs)
    sys.path.insert(0, work)
    return work


build_reference()

from splitfxm.domain import Domain  # noqa: E402
from splitfxm.simulation import Simulation  # noqa: E402
from splitfxm.schemes import default_scheme, FDSchemes, FVSchemes  # noqa: E402
from splitfxm.models.advection_diffusion import AdvectionDiffusion  # noqa: E402
from splitfxm.bc import apply_BC  # noqa: E402
from splitfxm.refine import Refiner  # noqa: E402
from splitfxm.constants import btype  # noqa: E402
from splitfxm.initialize import set_initial_condition  # noqa: E402

sys.path.insert(0, os.path.join(REF, "tutorials", "conservative_sod_shock"))
import euler1d as _cons  # noqa: E402
Euler1DConservative = _cons.Euler1DConservative
sys.path.pop(0)
del sys.modules["euler1d"]
sys.path.insert(0, os.path.join(REF, "tutorials", "sod_shock"))
import euler1d as _prim  # noqa: E402
Euler1D = _prim.Euler1D
sys.path.pop(0)

README_BCS = {
    "u": {"left": "periodic", "right": "periodic"},
    "v": {"left": {"dirichlet": 3}, "right": {"dirichlet": 4}},
    "w": {"left": {"dirichlet": 2}, "right": "periodic"},
}
NEUMANN0 = lambda names: {n: {"left": {"neumann": 0.}, "right": {"neumann": 0.}} for n in names}


ONLY = [t for t in os.environ.get("SFXM_GOLDEN_ONLY", "").split(",") if t]   # regenerate just these fixtures


def save(name, **kw):
    if ONLY and name not in ONLY:
        return
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **kw)
    print(f"wrote {name}.npz ({os.path.getsize(path)} B)")


def domain_state(d):
    return np.array(d.positions(), dtype=np.float64), np.array(d.values(), dtype=np.float64)


def coo_of(lil):
    c = lil.tocoo()
    order = np.lexsort((c.col, c.row))
    return c.row[order].astype(np.int64), c.col[order].astype(np.int64), c.data[order].astype(np.float64)


def bc_json(bcs):
    import json
    return json.dumps(bcs)


# ---------------------------------------------------------------------------------------------
# 1. Residual + Jacobian, advection-diffusion, every scheme the reference implements
#    (simulation.py:150-202 get_residuals_from_list, simulation.py:270-394 jacobian)
# ---------------------------------------------------------------------------------------------
def ad_case(tag, method, scheme, nx, nbl, nbr, ics, bcs, split, locs, c=0.2, nu=0.001, comps=("u", "v", "w"),
            perturb_seed=None, nonuniform=False):
    m = AdvectionDiffusion(c=c, nu=nu, method=method)
    d = Domain.from_size(nx, nbl, nbr, list(comps))
    s = Simulation(d, m, ics, bcs, scheme)
    if nonuniform:
        # a reference-made non-uniform mesh: two AMR passes of the reference's own Refiner
        s._r.set_criteria(2.0, 0.3, 0.3, -0.001)
        with contextlib.redirect_stdout(io.StringIO()):
            s._r.refine(d)
            s._r.refine(d)
        for cname, b in bcs.items():
            apply_BC(d, cname, b)
    if perturb_seed is not None:
        rng = np.random.default_rng(perturb_seed)
        for cell in d.interior():
            cell.set_values(cell.values() + 0.05 * rng.standard_normal(len(comps)))
        for cname, b in bcs.items():
            apply_BC(d, cname, b)
    x_all, U_all = domain_state(d)
    x0 = np.array(d.listify_interior(split, locs), dtype=np.float64)
    r = s.get_residuals_from_list(x0, split, locs)
    x_after, U_after = domain_state(d)
    J = s.jacobian(x0, split, locs)
    jr, jc, jv = coo_of(J)
    save(tag, what=f"AdvectionDiffusion {method} {scheme.name} nx={nx} nb=({nbl},{nbr}) split={split} locs={locs}: "
         "Simulation.get_residuals_from_list simulation.py:150-202; Simulation.jacobian simulation.py:270-394",
         model="ad", c=c, nu=nu, method=method, scheme=scheme.name, nx=d._nx, nbl=nbl, nbr=nbr,
         comps=np.array(comps), bcs=bc_json(bcs), split=split, locs=np.array(locs if locs else [], dtype=np.int64),
         x_all=x_all, U_all=U_all, x0=x0, residual=r, x_after=x_after, U_after=U_after,
         jac_row=jr, jac_col=jc, jac_val=jv, xmin=d.xmin(), xmax=d.xmax())


ICS3 = {"u": "gaussian", "v": "rarefaction", "w": "tophat"}
ICS2 = {"u": "gaussian", "v": "rarefaction"}
ad_case("ad_fvm_lf_nx20_split12", "FVM", FVSchemes.LAX_FRIEDRICHS, 20, 1, 1, ICS3, README_BCS, True, [1, 2])
ad_case("ad_fvm_lf_nx20_unsplit", "FVM", FVSchemes.LAX_FRIEDRICHS, 20, 1, 1, ICS3, README_BCS, False, None)
ad_case("ad_fdm_central_nx250_split1", "FDM", FDSchemes.CENTRAL, 250, 1, 1, ICS2, README_BCS, True, [1])
ad_case("ad_fdm_rbias_nx250_split1", "FDM", FDSchemes.RIGHT_BIAS, 250, 1, 2, ICS2, README_BCS, True, [1])
ad_case("ad_fdm_central_nx15_split12", "FDM", FDSchemes.CENTRAL, 15, 1, 1, ICS2, README_BCS, True, [1, 2])
ad_case("ad_fvm_central_nx33_unsplit", "FVM", FVSchemes.CENTRAL, 33, 1, 1, ICS3, README_BCS, False, None, perturb_seed=1)
ad_case("ad_fvm_upwind_nx33_split2", "FVM", FVSchemes.UPWIND, 33, 1, 1, ICS3, README_BCS, True, [2], c=-0.3,
        perturb_seed=2)
ad_case("ad_fvm_upwind_pos_nx20", "FVM", FVSchemes.UPWIND, 20, 1, 1, ICS3, README_BCS, False, None, c=0.1)
MIXED = {
    "u": {"left": "outflow", "right": "outflow"},
    "v": {"left": {"neumann": 0.5}, "right": {"neumann": -0.25}},
    "w": {"left": "periodic", "right": {"dirichlet": 1.5}},
}
ad_case("ad_fvm_lf_nx40_mixedbc", "FVM", FVSchemes.LAX_FRIEDRICHS, 40, 1, 1, ICS3, MIXED, False, None, perturb_seed=3)
ad_case("ad_fdm_rbias_nx40_mixedbc", "FDM", FDSchemes.RIGHT_BIAS, 40, 1, 2, ICS3, MIXED, True, [1, 2], perturb_seed=4)
ad_case("ad_fdm_rbias_nx40_nb21", "FDM", FDSchemes.RIGHT_BIAS, 40, 2, 1, ICS3, MIXED, False, None, perturb_seed=5)
ad_case("ad_fvm_lf_nonuniform", "FVM", FVSchemes.LAX_FRIEDRICHS, 30, 1, 1, ICS3, README_BCS, True, [1, 2],
        nonuniform=True)
ad_case("ad_fdm_central_nonuniform_mixed", "FDM", FDSchemes.CENTRAL, 30, 1, 1, ICS3, MIXED, False, None,
        nonuniform=True)
ad_case("ad_fdm_central_nx12_1var", "FDM", FDSchemes.CENTRAL, 12, 1, 1, {"u": "sine"},
        {"u": {"left": "periodic", "right": "periodic"}}, False, None, comps=("u",))
ad_case("ad_fvm_lf_nx16_2var", "FVM", FVSchemes.LAX_FRIEDRICHS, 16, 1, 1, {"u": "sine", "v": "tophat"},
        {"u": {"left": "periodic", "right": "periodic"}, "v": {"left": "outflow", "right": {"neumann": 1.0}}},
        True, [1], comps=("u", "v"))


# ---------------------------------------------------------------------------------------------
# 2. Euler (tutorials/*/euler1d.py) residual + Jacobian on the Sod state and on an evolved state
# ---------------------------------------------------------------------------------------------
def sod_case(tag, conservative, nx, evolve_t, split, locs):
    if conservative:
        names = ["rho", "rhou", "E"]
        m = Euler1DConservative(gamma=1.4, method="FVM")
    else:
        names = ["rho", "u", "p"]
        m = Euler1D(gamma=1.4, method="FVM")
    d = Domain.from_size(nx, 1, 1, names)
    bcs = NEUMANN0(names)
    s = Simulation(d, m, {}, bcs, default_scheme("FVM"))
    for cell in d.interior():
        left = cell.x() < 0.5
        cell.set_value(0, 1.0 if left else 0.125)
        cell.set_value(1, 0.0)
        if conservative:
            cell.set_value(2, 2.5 if left else 0.25)
        else:
            cell.set_value(2, 1.0 if left else 0.1)
    x_init, U_init = domain_state(d)
    nfev = [0]
    if evolve_t > 0:
        orig = s.get_residuals_from_list

        def counted(l, sp=False, lc=None):
            nfev[0] += 1
            return orig(l, sp, lc)
        s.get_residuals_from_list = counted
        s.evolve(split=split, split_locs=locs, t_diff=evolve_t)
        s.get_residuals_from_list = orig
    for cname, b in bcs.items():
        apply_BC(d, cname, b)
    x_all, U_all = domain_state(d)
    x0 = np.array(d.listify_interior(split, locs), dtype=np.float64)
    r = s.get_residuals_from_list(x0, split, locs)
    J = s.jacobian(x0, split, locs)
    jr, jc, jv = coo_of(J)
    save(tag, what=f"Euler {'conservative' if conservative else 'primitive'} FVM LF Sod nx={nx} after evolve({evolve_t}) RK45; "
         "tutorials/*/euler1d.py; simulation.py:150-202, 270-394, 396-472",
         model="euler_cons" if conservative else "euler_prim", gamma=1.4, method="FVM", scheme="LAX_FRIEDRICHS",
         nx=nx, nbl=1, nbr=1, comps=np.array(names), bcs=bc_json(bcs), split=split,
         locs=np.array(locs if locs else [], dtype=np.int64), x_init=x_init, U_init=U_init, evolve_t=evolve_t,
         nfev=nfev[0], x_all=x_all, U_all=U_all, x0=x0, residual=r, jac_row=jr, jac_col=jc, jac_val=jv,
         xmin=d.xmin(), xmax=d.xmax())


sod_case("sod_cons_nx100_t0", True, 100, 0.0, True, [1])
sod_case("sod_cons_nx100_t0p05", True, 100, 0.05, True, [1])
sod_case("sod_cons_nx60_unsplit_t0p02", True, 60, 0.02, False, None)
sod_case("sod_prim_nx100_t0", False, 100, 0.0, True, [1])
sod_case("sod_prim_nx60_t0p02", False, 60, 0.02, False, None)


# ---------------------------------------------------------------------------------------------
# 3. evolve (simulation.py:396-472): README driver RK45, explicit Euler, bounds
# ---------------------------------------------------------------------------------------------
def evolve_case(tag, method_name, t, max_step, nx=20, split=False, locs=None, fv=True, scheme=None):
    meth = "FVM" if fv else "FDM"
    m = AdvectionDiffusion(c=0.2, nu=0.001, method=meth)
    d = Domain.from_size(nx, 1, 1, ["u", "v", "w"])
    s = Simulation(d, m, ICS3, README_BCS, scheme or default_scheme(meth))
    x_init, U_init = domain_state(d)
    nfev = [0]
    orig = s.get_residuals_from_list

    def counted(l, sp=False, lc=None):
        nfev[0] += 1
        return orig(l, sp, lc)
    s.get_residuals_from_list = counted
    s.evolve(t, split, locs, method=method_name, max_step=max_step)
    x_all, U_all = domain_state(d)
    save(tag, what=f"AdvectionDiffusion {meth} nx={nx} evolve(t={t}, method={method_name}, max_step={max_step}) "
         "simulation.py:396-472 (README driver main.py:29)",
         model="ad", c=0.2, nu=0.001, method=meth, scheme=(scheme or default_scheme(meth)).name, nx=nx, nbl=1, nbr=1,
         comps=np.array(["u", "v", "w"]), bcs=bc_json(README_BCS), split=split,
         locs=np.array(locs if locs else [], dtype=np.int64), x_init=x_init, U_init=U_init, t=t,
         ode_method=method_name, max_step=max_step, nfev=nfev[0], x_all=x_all, U_all=U_all, xmin=0.0, xmax=1.0)


evolve_case("evolve_readme_rk45", "RK45", 0.1, np.inf)
evolve_case("evolve_readme_rk45_split", "RK45", 0.25, np.inf, split=True, locs=[1, 2])
evolve_case("evolve_rk45_maxstep", "RK45", 0.3, 0.05, nx=40, fv=False)
evolve_case("evolve_readme_radau", "Radau", 0.1, np.inf)
evolve_case("evolve_readme_bdf_split", "BDF", 0.2, np.inf, split=True, locs=[1, 2], fv=False)
evolve_case("evolve_readme_radau_split", "Radau", 0.2, np.inf, split=True, locs=[1, 2])
evolve_case("evolve_radau_split1_fd", "Radau", 0.2, np.inf, nx=24, split=True, locs=[1], fv=False)
evolve_case("evolve_readme_rk23", "RK23", 0.1, np.inf)
evolve_case("evolve_euler", "Euler", 1.0, 0.03, nx=30)
evolve_case("evolve_euler_split_fd", "Euler", 0.5, 0.01, nx=25, split=True, locs=[1], fv=False)


# ---------------------------------------------------------------------------------------------
# 4. AMR (refine.py:90-282): seeded random meshes + the evolving Sod state
# ---------------------------------------------------------------------------------------------
def refine_cases():
    rng = np.random.default_rng(20261019)
    xs, Us, outs_x, outs_U, crit, offs_in, offs_out, nvs = [], [], [], [], [], [0], [0], []
    n_cases = 120
    for k in range(n_cases):
        n = int(rng.integers(6, 70))
        nv = int(rng.integers(1, 4))
        names = ["a", "b", "c"][:nv]
        d = Domain.from_size(n, 1, 1, names)
        cells = d.interior()
        if k % 2 == 1:
            # non-uniform but sorted, distinct positions
            xi = np.sort(rng.random(n))
            xi[0], xi[-1] = 0.0, 1.0
            for cell, xv in zip(cells, xi):
                cell._x = float(xv)
        kind = k % 4
        for cell in cells:
            x = cell.x()
            vals = []
            for j in range(nv):
                if kind == 0:
                    vals.append(np.tanh(20 * (x - 0.3 - 0.2 * j)))
                elif kind == 1:
                    vals.append(rng.standard_normal())
                elif kind == 2:
                    vals.append(1.0 if 0.3 < x < 0.6 else 0.0)
                else:
                    vals.append(np.exp(-200 * (x - 0.25) ** 2) + 0.01 * j * x)
            cell.set_values(np.array(vals, dtype=np.float64))
        ratio = float(rng.choice([2.0, 3.0, 10.0]))
        slope = float(rng.choice([0.2, 0.5, 0.8]))
        curve = float(rng.choice([0.2, 0.5, 0.8]))
        prune = float(rng.choice([-0.001, 0.0, 0.05, 0.1, 0.2]))
        r = Refiner()
        r.set_criteria(ratio, slope, curve, prune)
        xin = np.array(d.positions(interior=True))
        Uin = np.array(d.values(interior=True))
        with contextlib.redirect_stdout(io.StringIO()):
            r.refine(d)
        xout = np.array(d.positions(interior=True))
        Uout = np.array(d.values(interior=True))
        xs.append(xin); Us.append(Uin.ravel()); outs_x.append(xout); outs_U.append(Uout.ravel())
        crit.append([ratio, slope, curve, prune]); nvs.append(nv)
        offs_in.append(offs_in[-1] + len(xin)); offs_out.append(offs_out[-1] + len(xout))
    save("refine_random", what="Refiner.refine refine.py:90-282 on 120 seeded meshes (n 6..69, nv 1..3, uniform and "
         "non-uniform, tanh/noise/tophat/gaussian fields); arrays are concatenated, offsets delimit the cases",
         x_in=np.concatenate(xs), U_in=np.concatenate(Us), x_out=np.concatenate(outs_x), U_out=np.concatenate(outs_U),
         criteria=np.array(crit), nv=np.array(nvs), off_in=np.array(offs_in), off_out=np.array(offs_out))


refine_cases()


def refine_sod():
    names = ["rho", "rhou", "E"]
    m = Euler1DConservative(gamma=1.4, method="FVM")
    d = Domain.from_size(100, 1, 1, names)
    bcs = NEUMANN0(names)
    s = Simulation(d, m, {}, bcs, default_scheme("FVM"))
    for cell in d.interior():
        left = cell.x() < 0.5
        cell.set_value(0, 1.0 if left else 0.125)
        cell.set_value(1, 0.0)
        cell.set_value(2, 2.5 if left else 0.25)
    s._r.set_criteria(2.0, 0.05, 0.05, -0.001)
    xs, Us, offs = [], [], [0]
    for p in range(4):
        s.evolve(split=False, t_diff=0.01)
        xs.append(np.array(d.positions(interior=True))); Us.append(np.array(d.values(interior=True)).ravel())
        offs.append(offs[-1] + len(xs[-1]))
        with contextlib.redirect_stdout(io.StringIO()):
            s._r.refine(d)
        xs.append(np.array(d.positions(interior=True))); Us.append(np.array(d.values(interior=True)).ravel())
        offs.append(offs[-1] + len(xs[-1]))
    save("refine_sod", what="conservative Sod nx=100: 4 x [evolve(0.01) RK45 -> Refiner.refine] with "
         "set_criteria(2.0,0.05,0.05,-0.001); states stored before and after each refine "
         "(tutorials/conservative_sod_shock, refine.py:90-282, simulation.py:396-472)",
         x=np.concatenate(xs), U=np.concatenate(Us), off=np.array(offs), criteria=np.array([2.0, 0.05, 0.05, -0.001]))


refine_sod()


# ---------------------------------------------------------------------------------------------
# 5. Initial conditions (initialize.py:7-59) and from_size ghosts (domain.py:82-97)
# ---------------------------------------------------------------------------------------------
def ic_cases():
    out = {}
    for nx, xmin, xmax in ((20, 0.0, 1.0), (57, -1.0, 3.0)):
        d = Domain.from_size(nx, 2, 2, ["a", "b", "c", "d"], xmin, xmax)
        for name, kind in zip(["a", "b", "c", "d"], ["tophat", "sine", "rarefaction", "gaussian"]):
            set_initial_condition(d, name, kind)
        x_all, U_all = domain_state(d)
        out[f"x_{nx}"] = x_all
        out[f"U_{nx}"] = U_all
    save("initial_conditions", what="Domain.from_size domain.py:48-99 + set_initial_condition initialize.py:7-59 "
         "(tophat, sine, rarefaction, gaussian) on [0,1] nx=20 and [-1,3] nx=57, 2+2 ghosts", **out)


ic_cases()
print("done")
