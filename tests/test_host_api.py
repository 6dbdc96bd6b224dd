"""Host-side mirror of the reference API (no GPU): the C ABI exports every declared symbol; Domain / Cell / Boundary
views write through; boundary conditions, periodic-band helpers, bounds tiling and the error behaviour follow the
reference (splitfxm/domain.py, cell.py, boundary.py, bc.py, simulation.py:204-264)."""
import copy
import os
import re

import numpy as np
import pytest

import splitfxm_b200 as sf
from splitfxm_b200 import _lib
from splitfxm_b200.bc import apply_BC, extend_band, get_periodic_bcs
from splitfxm_b200.constants import btype
from splitfxm_b200.models.advection_diffusion import AdvectionDiffusion

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "sfxm_b200.h")).read()
    declared = set(re.findall(r"\b(sfxm_[a-z0-9_]+)\s*\(", header))
    declared -= {"sfxm_problem"}
    assert len(declared) >= 20
    lib = _lib.load()
    missing = [name for name in sorted(declared) if not hasattr(lib, name)]
    assert not missing, missing
    assert set(_lib.SIGNATURES) == declared          # the ctypes table binds exactly the declared entry points
    assert lib.sfxm_abi_version() == 1


def test_problem_struct_layout_matches_header():
    """ctypes mirror of struct sfxm_problem: field order/size (a mismatch would silently corrupt kernel arguments)."""
    assert _lib.Problem.model.offset == 0 and _lib.Problem.nx.offset == 24 and _lib.Problem.batch.offset == 32
    assert _lib.Problem.bc_lval.offset == _lib.Problem.bc_rtype.offset + 16
    assert ctypes_sizeof(_lib.Problem) % 8 == 0


def ctypes_sizeof(t):
    import ctypes
    return ctypes.sizeof(t)


def test_hot_path_raises_without_gpu_instead_of_falling_back():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    d = sf.Domain.from_size(20, 1, 1, ["u"])
    s = sf.Simulation(d, AdvectionDiffusion(0.2, 0.001, "FVM"), {"u": "gaussian"},
                      {"u": {"left": "periodic", "right": "periodic"}}, sf.default_scheme("FVM"))
    with pytest.raises(sf.SFXM):
        s.get_residuals_from_list(d.listify_interior(False))
    with pytest.raises(sf.SFXM):
        s.evolve(0.1)
    with pytest.raises(sf.SFXM):
        s._r.refine(d)


def test_domain_accessors_and_views():
    d = sf.Domain.from_size(5, 1, 2, ["a", "b"], 0.0, 2.0)
    assert d.ilo() == 1 and d.ihi() == 5 and d.nb(btype.LEFT) == 1 and d.nb(btype.RIGHT) == 2
    assert d.num_components() == 2 and d.component_index("b") == 1 and d.component_name(0) == "a"
    assert np.allclose(d.positions(interior=True), np.linspace(0, 2, 5))
    assert len(d.cells()) == 8 and len(d.interior()) == 5
    lb, rb = d.boundaries()
    assert len(lb) == 1 and len(rb) == 2 and lb[0].x() < 0 and rb[1].x() > rb[0].x() > 2.0
    # cell views write through to the arrays the device path uploads
    c = d.interior()[2]
    c.set_value(1, 7.5)
    assert d.values(interior=True)[2, 1] == 7.5
    c.set_values(np.array([1.0, 2.0]))
    assert np.array_equal(d.interior_U()[2], [1.0, 2.0])
    c._x = 1.01
    assert d.positions(interior=True)[2] == 1.01
    c.update(0.5, np.array([2.0, 2.0]))
    assert np.array_equal(d.values()[3], [2.0, 3.0])
    assert d.interior()[0] < d.interior()[1] and d.interior()[1] == d.interior()[1]
    flat = d.listify_interior(True, [1])
    assert np.array_equal(flat[:5], d.values(interior=True)[:, 0])
    with pytest.raises(sf.SFXM):
        d.listify_interior(True, None)
    with pytest.raises(sf.SFXM):
        d.listify_interior(True, [3])


def test_domain_from_objects_deepcopy_set_interior():
    cells = [sf.Cell(x, np.array([x, 2 * x])) for x in (0.0, 0.5, 1.0)]
    lb = [sf.Boundary(-0.5, btype.LEFT, np.zeros(2))]
    rb = [sf.Boundary(1.5, btype.RIGHT, np.zeros(2))]
    d = sf.Domain(cells, lb, rb, ["p", "q"])
    assert d.positions() == [-0.5, 0.0, 0.5, 1.0, 1.5]
    cells[1].set_value(0, 9.0)                      # the original objects stay bound to the domain
    assert d.values()[2, 0] == 9.0
    d2 = copy.deepcopy(d)
    d2.interior()[0].set_value(0, -1.0)
    assert d.values()[1, 0] == 0.0 and d2.values()[1, 0] == -1.0
    new = d.interior() + [sf.Cell(0.75, np.array([1.0, 1.0]))]
    new.sort()
    d.set_interior(new)
    assert d.positions(interior=True) == [0.0, 0.5, 0.75, 1.0] and d.ihi() == 4
    with pytest.raises(sf.SFXM):
        sf.Boundary(0.5, btype.RIGHT, np.zeros(1), xmin=1.0, xmax=2.0)
    with pytest.raises(sf.SFXM):
        rb[0].set_x(0.5)


def _line_domain(nbl=1, nbr=1):
    d = sf.Domain.from_size(6, nbl, nbr, ["u"])
    for k, c in enumerate(d.interior()):
        c.set_value(0, float(k * k))
    return d


def test_boundary_conditions_exact_values():
    d = _line_domain(2, 2)
    apply_BC(d, "u", {"left": "periodic", "right": "periodic"})
    v = d.values()[:, 0]
    assert list(v[:2]) == [16.0, 25.0] and list(v[-2:]) == [0.0, 1.0]
    x = np.array(d.positions())
    assert np.isclose(x[1], -0.2) and np.isclose(x[0], -0.4) and np.isclose(x[-2], 1.2) and np.isclose(x[-1], 1.4)
    apply_BC(d, "u", {"left": {"dirichlet": 3}, "right": {"dirichlet": -2.5}})
    assert list(d.values()[:2, 0]) == [3.0, 3.0] and list(d.values()[-2:, 0]) == [-2.5, -2.5]
    apply_BC(d, "u", {"left": {"neumann": 2.0}, "right": {"neumann": 2.0}})
    v = d.values()[:, 0]
    assert np.isclose(v[1], 0.0 - 2.0 * 0.2) and np.isclose(v[0], v[1] - 2.0 * 0.2)
    assert np.isclose(v[-2], 25.0 + 2.0 * 0.2) and np.isclose(v[-1], v[-2] + 2.0 * 0.2)
    apply_BC(d, "u", {"left": "outflow", "right": "outflow"})
    v = d.values()[:, 0]
    assert np.isclose(v[1], 0.0 - (1.0 / 0.2) * 0.2) and np.isclose(v[-2], 25.0 + (9.0 / 0.2) * 0.2)
    for bad in ({"left": "periodic"}, {"left": "periodic", "right": "reflect"},
                {"left": {"dirichlet": "x"}, "right": "periodic"}, {"left": {"robin": 1.0}, "right": "periodic"}):
        with pytest.raises(sf.SFXM):
            apply_BC(d, "u", bad)


def test_periodic_helpers():
    d = sf.Domain.from_size(10, 1, 2, ["u", "v", "w"])
    bcs = {"u": {"left": "periodic", "right": "periodic"}, "v": {"left": {"dirichlet": 1}, "right": {"dirichlet": 2}},
           "w": {"left": {"dirichlet": 2}, "right": "periodic"}}
    assert get_periodic_bcs(bcs, d) == {0: ["left", "right"], 2: ["right"]}
    ilo, ihi = d.ilo(), d.ihi()
    assert sorted(extend_band([ihi - 1, ihi], ["right"], ihi, d)) == [ilo, ilo + 1, ihi - 1, ihi]
    assert sorted(extend_band([ilo, ilo + 1, ilo + 2], ["left"], ilo, d)) == [ilo, ilo + 1, ilo + 2, ihi]
    assert sorted(extend_band([5, 6], ["left", "right"], 5, d)) == [5, 6]
    with pytest.raises(sf.SFXM):
        extend_band([1], ["up"], 1, d)


def test_simulation_host_logic():
    comps = ["u", "v", "w"]
    bcs = {"u": {"left": "periodic", "right": "periodic"}, "v": {"left": {"dirichlet": 3}, "right": {"dirichlet": 4}},
           "w": {"left": {"dirichlet": 2}, "right": "periodic"}}
    d = sf.Domain.from_size(4, 1, 1, comps)
    s = sf.Simulation(d, AdvectionDiffusion(0.2, 0.001, "FVM"), {"u": "gaussian", "v": "rarefaction", "w": "tophat"},
                      bcs, sf.default_scheme("FVM"))
    assert s.get_shape_from_list(list(range(12))) == (4, 3)
    with pytest.raises(sf.SFXM):
        s.get_shape_from_list(list(range(11)))
    assert s.extend_bounds(None, 4, 3) is None
    assert s.extend_bounds([[0, 1, 2], [5, 6, 7]], 2, 3) == [[0, 1, 2, 0, 1, 2], [5, 6, 7, 5, 6, 7]]
    assert s.extend_bounds([[0, 1, 2], [5, 6, 7]], 2, 3, True, [1]) == [[0, 0, 1, 2, 1, 2], [5, 5, 6, 7, 6, 7]]
    for bad in ([[0, 1, 2]], [[0, 1], [5, 6, 7]]):
        with pytest.raises(sf.SFXM):
            s.extend_bounds(bad, 2, 3)
    with pytest.raises(sf.SFXM):
        s.extend_bounds([[0, 1, 2], [5, 6, 7]], 2, 3, True, None)
    s.initialize_from_list(np.arange(12.0), True, [1, 2])
    assert np.array_equal(d.values(interior=True)[:, 0], [0, 1, 2, 3]) and np.array_equal(d.values(interior=True)[:, 2], [8, 9, 10, 11])
    with pytest.raises(sf.SFXM):
        s.initialize_from_list(np.arange(12.0), True, None)
    # stencil size must match the ghost cells (simulation.py:73-75); names-only schemes are rejected
    with pytest.raises(sf.SFXM):
        sf.Simulation(sf.Domain.from_size(4, 1, 2, comps), AdvectionDiffusion(0.2, 0.001, "FVM"), {}, bcs, sf.FVSchemes.UPWIND)
    with pytest.raises(sf.SFXM):
        sf.Simulation(sf.Domain.from_size(4, 1, 1, comps), AdvectionDiffusion(0.2, 0.001, "FVM"), {}, bcs, sf.FVSchemes.WENO)
    with pytest.raises(NotImplementedError):
        sf.Simulation(sf.Domain.from_size(4, 1, 1, comps), AdvectionDiffusion(0.2, 0.001, "FVM"), {"u": "bump"}, bcs,
                      sf.default_scheme("FVM"))
    with pytest.raises(sf.SFXM):
        AdvectionDiffusion(0.2, 0.001, "FEM")
    with pytest.raises(sf.SFXM):
        sf.default_scheme("FEM")


def test_refiner_criteria_validation():
    r = sf.Refiner()
    assert (r._ratio, r._slope, r._curve, r._prune, r._npmax) == (10.0, 0.8, 0.8, -0.001, 1000)
    for args in ((1.5, 0.5, 0.5, 0.1), (3.0, 1.5, 0.5, 0.1), (3.0, 0.5, -0.1, 0.1), (3.0, 0.5, 0.5, 0.6)):
        with pytest.raises(sf.SFXM):
            r.set_criteria(*args)
    r.set_criteria(3.0, 0.5, 0.4, 0.1)
    r.set_max_points(50)
    assert (r._ratio, r._slope, r._curve, r._prune, r._npmax) == (3.0, 0.5, 0.4, 0.1, 50)
