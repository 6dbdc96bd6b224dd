"""AMR on a slab-decomposed mesh (splitfxm_b200/slab_refine.py): G slabs as G threads on one GPU must reproduce the
UNMODIFIED reference Refiner's meshes (tests/golden/refine_random.npz, refine_sod.npz) bit for bit, with and without
the rebalancing step, and print the reference's report with global indices."""
import contextlib
import io

import numpy as np
import pytest

from helpers import load

pytestmark = pytest.mark.gpu


def _refine_on_slabs(x_in, U_in, criteria, G, rebalance, report=False, npmax=None):
    import torch
    from splitfxm_b200.distributed import slab_bounds
    from splitfxm_b200.slab_refine import SlabRefiner
    from splitfxm_b200.spike import ThreadGroup
    n = x_in.shape[0]

    def body(rank, comm):
        lo, hi = slab_bounds(n, G, rank)
        r = SlabRefiner()
        r.set_criteria(*criteria)
        if npmax is not None:
            r.set_max_points(npmax)
        x = torch.from_numpy(x_in[lo:hi].copy()).cuda()
        U = torch.from_numpy(U_in[lo:hi].copy()).cuda()
        x_new, U_new, lo_new, total = r.refine_slab(x, U, lo, n, comm, rebalance=rebalance, report=report)
        return x_new.cpu().numpy(), U_new.cpu().numpy(), lo_new, total

    out = ThreadGroup(G).run(body)
    total = out[0][3]
    pos = 0
    for x_new, _, lo_new, tot in out:                      # consecutive, gap-free ranges in rank order
        assert lo_new == pos and tot == total
        pos += x_new.shape[0]
    assert pos == total
    if rebalance:
        sizes = [o[0].shape[0] for o in out]
        assert max(sizes) - min(sizes) <= 1
    return np.concatenate([o[0] for o in out]), np.concatenate([o[1] for o in out])


@pytest.mark.parametrize("G,rebalance", [(2, False), (3, True), (4, True)])
def test_slab_refine_random_meshes_identical_to_reference(G, rebalance):
    z = load("refine_random")
    n_cases = z["nv"].shape[0]
    done = grew = shrank = 0
    for k in range(n_cases):
        nv = int(z["nv"][k])
        a, b = int(z["off_in"][k]), int(z["off_in"][k + 1])
        c, d = int(z["off_out"][k]), int(z["off_out"][k + 1])
        if (b - a) // G < 6:
            continue                                       # slabs must hold the six-cell overlap
        x_in = z["x_in"][a:b]
        ua = int(np.sum((z["off_in"][1:k + 1] - z["off_in"][:k]) * z["nv"][:k]))
        U_in = z["U_in"][ua:ua + (b - a) * nv].reshape(b - a, nv)
        uc = int(np.sum((z["off_out"][1:k + 1] - z["off_out"][:k]) * z["nv"][:k]))
        U_ref = z["U_out"][uc:uc + (d - c) * nv].reshape(d - c, nv)
        x_ref = z["x_out"][c:d]
        x_new, U_new = _refine_on_slabs(x_in, U_in, [float(v) for v in z["criteria"][k]], G, rebalance)
        assert x_new.shape == x_ref.shape, f"case {k}: {x_new.shape[0]} cells, reference {x_ref.shape[0]}"
        assert np.array_equal(x_new, x_ref), f"case {k}: positions differ"
        assert np.array_equal(U_new, U_ref), f"case {k}: values differ"
        done += 1
        grew += x_ref.shape[0] > x_in.shape[0]
        shrank += np.setdiff1d(x_in, x_ref).size > 0
    assert done >= 60 and grew > 25 and shrank > 10


def test_slab_refine_sod_sequence():
    z = load("refine_sod")
    off = z["off"]
    crit = [float(v) for v in z["criteria"]]
    for p in range(4):
        a, b, c = int(off[2 * p]), int(off[2 * p + 1]), int(off[2 * p + 2])
        x_in, U_in = z["x"][a:b], z["U"][3 * a:3 * b].reshape(-1, 3)
        x_ref, U_ref = z["x"][b:c], z["U"][3 * b:3 * c].reshape(-1, 3)
        x_new, U_new = _refine_on_slabs(x_in, U_in, crit, 4, True)
        assert np.array_equal(x_new, x_ref) and np.array_equal(U_new, U_ref)


def test_slab_refine_long_deletion_run_crosses_slabs():
    """A flat field with pruning on: one deletion run spans every slab, so the odd/even rule depends on the run start
    carried across ranks; the report lists the same global indices as the single-GPU Refiner."""
    import torch
    from splitfxm_b200.refine import Refiner
    n = 203
    x = np.linspace(0.0, 1.0, n)
    U = np.stack([np.where(x < 0.1, 1.0 - 5 * x, 0.5), np.tanh(30 * (x - 0.8))], axis=1)
    crit = (2.0, 0.3, 0.3, 0.05)
    r = Refiner()
    r.set_criteria(*crit)
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        x_ref, U_ref, flags = r.refine_arrays(x, U)
        r.show_changes(*r._dicts_from_flags({k: torch.from_numpy(v) for k, v in flags.items()}, ["component 0", "component 1"]))
    assert (flags["keep"] == 0).sum() > 40                 # a long run of deletions exists
    for G in (2, 5):
        out = io.StringIO()
        with contextlib.redirect_stdout(out):
            x_new, U_new = _refine_on_slabs(x, U, crit, G, True, report=True)
        assert np.array_equal(x_new, x_ref) and np.array_equal(U_new, U_ref)
        assert out.getvalue() == buf.getvalue()


def test_slab_refine_max_points_and_large_mesh():
    import splitfxm_b200 as sf
    x = np.linspace(0.0, 1.0, 100)
    U = np.sin(6 * x)[:, None]
    with pytest.raises(sf.SFXM):
        _refine_on_slabs(x, U, (2.0, 0.2, 0.2, -0.001), 2, True, npmax=50)
    # 1e6 well-resolved cells on 4 slabs: nothing changes (size-independent property)
    n = 1_000_000
    x = np.linspace(0.0, 1.0, n)
    U = np.stack([np.sin(2 * np.pi * x), x ** 2, 1.0 + 0 * x], axis=1)
    x_new, U_new = _refine_on_slabs(x, U, (10.0, 0.8, 0.8, -0.001), 4, True, npmax=2 * n)
    assert np.array_equal(x_new, x) and np.array_equal(U_new, U)
