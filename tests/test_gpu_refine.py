"""AMR kernels (csrc/sfxm_amr.cu) against meshes produced by the UNMODIFIED reference Refiner
(tests/golden/refine_random.npz: 120 seeded meshes; refine_sod.npz: the evolving Sod tube).  Mesh indexing is
integer work: positions and values must be bit-identical."""
import contextlib
import io

import numpy as np
import pytest

from helpers import load

pytestmark = pytest.mark.gpu


def test_random_meshes_identical_to_reference():
    from splitfxm_b200.refine import Refiner
    z = load("refine_random")
    n_cases = z["nv"].shape[0]
    grew = shrank = 0
    for k in range(n_cases):
        nv = int(z["nv"][k])
        a, b = int(z["off_in"][k]), int(z["off_in"][k + 1])
        c, d = int(z["off_out"][k]), int(z["off_out"][k + 1])
        x_in = z["x_in"][a:b]
        # U_in is the concatenation of the raveled [n, nv] blocks
        ua = int(np.sum((z["off_in"][1:k + 1] - z["off_in"][:k]) * z["nv"][:k]))
        U_in = z["U_in"][ua:ua + (b - a) * nv].reshape(b - a, nv)
        uc = int(np.sum((z["off_out"][1:k + 1] - z["off_out"][:k]) * z["nv"][:k]))
        U_ref = z["U_out"][uc:uc + (d - c) * nv].reshape(d - c, nv)
        x_ref = z["x_out"][c:d]
        r = Refiner()
        r.set_criteria(*[float(v) for v in z["criteria"][k]])
        x_new, U_new, _ = r.refine_arrays(x_in, U_in)
        assert x_new.shape == x_ref.shape, f"case {k}: {x_new.shape[0]} cells, reference {x_ref.shape[0]}"
        assert np.array_equal(x_new, x_ref), f"case {k}: positions differ"
        assert np.array_equal(U_new, U_ref), f"case {k}: values differ"
        grew += x_ref.shape[0] > x_in.shape[0]
        shrank += np.setdiff1d(x_in, x_ref).size > 0
    assert grew > 50 and shrank > 20     # the fixture exercises both insertion and deletion


def test_sod_refinement_sequence():
    """States before/after each of the reference's four refine passes on the evolving conservative Sod tube."""
    from splitfxm_b200.refine import Refiner
    z = load("refine_sod")
    off = z["off"]
    r = Refiner()
    r.set_criteria(*[float(v) for v in z["criteria"]])
    for p in range(4):
        a, b, c = int(off[2 * p]), int(off[2 * p + 1]), int(off[2 * p + 2])
        x_in, U_in = z["x"][a:b], z["U"][3 * a:3 * b].reshape(-1, 3)
        x_ref, U_ref = z["x"][b:c], z["U"][3 * b:3 * c].reshape(-1, 3)
        x_new, U_new, _ = r.refine_arrays(x_in, U_in)
        assert np.array_equal(x_new, x_ref) and np.array_equal(U_new, U_ref)


def test_refine_through_domain_api_and_report():
    import splitfxm_b200 as sf
    d = sf.Domain.from_size(30, 1, 1, ["u"])
    for cell in d.interior():
        cell.set_value(0, np.tanh(40 * (cell.x() - 0.5)))
    r = sf.Refiner()
    r.set_criteria(2.0, 0.2, 0.2, -0.001)
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        r.refine(d)
    out = buf.getvalue()
    assert "Refining grid..." in out and "No new points deleted" in out
    assert d._nx > 30 and len(d.interior()) == d._nx
    xs = np.array(d.positions(interior=True))
    assert np.all(np.diff(xs) > 0)
    r.set_max_points(10)
    with pytest.raises(sf.SFXM):
        r.refine(d)


def test_large_mesh_idempotent_when_resolved():
    """1e6 cells of a smooth, well-resolved field: nothing to insert or delete (size-independent property)."""
    from splitfxm_b200.refine import Refiner
    n = 1_000_000
    x = np.linspace(0.0, 1.0, n)
    U = np.stack([np.sin(2 * np.pi * x), x ** 2, 1.0 + 0 * x], axis=1)
    r = Refiner()
    r.set_max_points(2 * n)
    x_new, U_new, flags = r.refine_arrays(x, U)
    assert np.array_equal(x_new, x) and np.array_equal(U_new, U)
    assert flags["loc"].sum() == 0 and (flags["keep"] == 1).all()


def test_refinement_at_config3_scale_identical_to_reference():
    """BASELINE config 3's mesh size: three consecutive passes on a 100,000-cell Sod-like state (-> 98,683 -> 119,221
    -> 129,044 cells, each pass inserting ~2e4 and deleting 5e3-2e4 cells) through the Domain API, device-resident
    between the passes.  Identical meshes are identical byte strings: sizes, SHA-256 digests of positions and values and
    every 997th cell are compared with what the unmodified reference produced (tests/golden/make_amr_large_golden.py)."""
    import os
    import sys
    import splitfxm_b200 as sf
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
    import make_amr_large_golden as g
    z = load("amr_large")
    x, U = g.sod_like_state(int(z["nx0"]))
    assert g.digest(x) == str(z["input_x_sha"]) and g.digest(U) == str(z["input_U_sha"])     # same input on this machine
    d = sf.Domain.from_size(int(z["nx0"]), 1, 1, ["rho", "rhou", "E"])
    assert np.array_equal(d.interior_x(), x)
    d.interior_U()[:, :] = U
    r = sf.Refiner()
    r.set_criteria(*[float(v) for v in z["criteria"]])
    r.set_max_points(10 * int(z["nx0"]))
    for k in range(g.PASSES):
        with contextlib.redirect_stdout(io.StringIO()):
            r.refine(d)
        assert not d._host_ok                       # the new mesh stays on the device between the passes
        dx, dU = d.device_arrays()
        xs, Us = dx[1:-1].cpu().numpy(), dU[1:-1].cpu().numpy()
        assert xs.shape[0] == int(z[f"n{k}"]), (k, xs.shape[0])
        assert np.array_equal(xs[::997], z[f"x_sample{k}"]) and np.array_equal(Us[::997], z[f"U_sample{k}"])
        assert g.digest(xs) == str(z[f"x_sha{k}"]) and g.digest(Us) == str(z[f"U_sha{k}"])
