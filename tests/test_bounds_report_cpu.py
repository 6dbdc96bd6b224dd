"""Bound warnings of evolve (reference: apply_bounds, simhelper.py:91-116): one line per offending unknown in index order,
from a host copy (report_bounds) and from a tensor where the state lives (DeviceBounds, here on CPU tensors)."""
import contextlib
import io

import numpy as np
import torch


def _expected(y, lower, upper, t):
    lines = []
    for idx, (val, lb, ub) in enumerate(zip(y, lower, upper)):          # the reference's loop, restated
        if val < lb:
            lines.append(f"Warning: At t={t}, bounds exceeded for variable index {idx}, value = {val}, lower bound = {lb}")
        elif val > ub:
            lines.append(f"Warning: At t={t}, bounds exceeded for variable index {idx}, value = {val}, upper bound = {ub}")
    return "".join(l + "\n" for l in lines)


def test_reports_match_the_reference_loop():
    from splitfxm_b200.simhelper import DeviceBounds, report_bounds
    rng = np.random.default_rng(0)
    y = rng.standard_normal(40)
    lower, upper = np.full(40, -0.8), np.full(40, 0.9)
    upper[3] = -5.0
    want = _expected(y, lower, upper, 0.25)
    assert want.count("lower bound") > 2 and want.count("upper bound") > 2
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        report_bounds(y, [lower, upper], 0.25)
    assert buf.getvalue() == want
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        n = DeviceBounds([lower.tolist(), upper.tolist()], "cpu").report(torch.from_numpy(y).reshape(-1, 1), 0.25)
    assert buf.getvalue() == want and n == want.count("\n")


def test_nothing_is_printed_inside_the_bounds():
    from splitfxm_b200.simhelper import DeviceBounds
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        n = DeviceBounds([[-1.0] * 5, [1.0] * 5], "cpu").report(torch.zeros(5, dtype=torch.float64), 1.0)
    assert n == 0 and buf.getvalue() == ""
