"""bench.py's reference arm runs on host cores only: check its one JSON line against the contract here on CPU, and that
the oracle port it times agrees with the committed reference fixture for the same shape."""
import json
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_contract_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "0", "--ref-members", "32", "--newton-steps", "1"], capture_output=True, text=True,
                         timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip().startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better",
                "scaling", "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in d, key
    assert d["impl"] == "reference" and d["metric"] == "jacobian_assemblies_per_s" and d["dtype"] == "f64"
    assert d["value"] > 0 and d["steps"] == 1 and d["higher_is_better"] is True and d["vs_baseline"] is None
    assert "workload" in d["config"]
    cb = d["cpu_baseline"]
    sys.path.insert(0, ROOT)
    from baseline import reference_env
    # the unmodified reference (baseline/_ref, git-ignored, installed by __graft_entry__.build()) when present, else the port
    assert cb["kind"] == ("reference" if reference_env.available() else "port")
    assert cb["cores"] >= 1 and cb["value"] == d["value"] and "sample" in cb
    if cb["kind"] == "reference":
        nw = d["newton"]
        assert nw["metric"] == "newton_solves_per_s" and nw["value"] > 0 and nw["iterations_min_max"][0] >= 1
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_both_arms_share_one_config():
    """The driver compares the two arms on `config`: one function builds it for both."""
    sys.path.insert(0, ROOT)
    import bench
    assert bench.bench_config(65536, 8) == bench.bench_config(65536, 8)
    assert "members_per_gpu" in bench.bench_config(128, 2) and "workload" in bench.bench_config(128, 2)


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1",
                          "--warmup", "0"], capture_output=True, text=True, timeout=120, cwd=ROOT, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_cpu_arm_assembles_the_fixture_jacobian():
    """The banded Jacobian the CPU arm times (oracle/core.py jacobian_banded) holds the entries of the reference's
    sparse matrix for the nx = 250 benchmark shape (tests/golden/ad_fdm_central_nx250_split1.npz)."""
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from helpers import load, oracle_problem
    from oracle import core
    z = load("ad_fdm_central_nx250_split1")
    P, m = oracle_problem(z)
    U = core.to_block(z["x0"], m["nx"], 3, m["split"], m["locs"])
    Jb = P.jacobian_banded(U)                                     # [nx, W, nv, nv]
    rows, cols, vals = P.jacobian_coo(z["x0"], m["split"], m["locs"])
    assert np.array_equal(rows, z["jac_row"]) and np.array_equal(cols, z["jac_col"]) and np.array_equal(vals, z["jac_val"])
    # every stored value appears in the band at (cell of the row, slot of the column, variable pair)
    idx = core.flat_index(m["nx"], 3, m["split"], m["locs"])      # [nx, nv] -> flat position
    inv = {int(idx[p, j]): (p, j) for p in range(m["nx"]) for j in range(3)}
    for r, c, v in list(zip(rows, cols, vals))[::37]:
        (pi, ri), (pc, cj) = inv[int(r)], inv[int(c)]
        d = (pc - pi + 1) % m["nx"]
        assert d in (0, 1, 2) and Jb[pi, d, ri, cj] == v


def test_gpu_arm_has_no_cpu_fallback():
    """Without a CUDA device the product arm must fail loudly - no JSON line, non-zero exit (there is no CPU path)."""
    import torch
    if torch.cuda.is_available():
        import pytest
        pytest.skip("a CUDA device is present")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1", "--warmup", "0", "--no-extra"],
                         capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert out.returncode != 0
    assert not [l for l in out.stdout.splitlines() if l.strip().startswith("{")]
