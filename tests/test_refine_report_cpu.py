"""The stdout report of a refinement pass (splitfxm_b200/refine.py show_changes) in the reference's format
(refine.py:284-312: banner, inserted indices, reasons, deleted indices), from decision flags alone - host logic, no GPU."""
import contextlib
import io

import numpy as np
import pytest


def _report(flags, names):
    from splitfxm_b200.refine import Refiner
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        Refiner().show_changes(flags, names)
    return buf.getvalue()


def test_report_lists_insertions_reasons_and_deletions():
    n = 12
    loc = np.zeros(n, dtype=np.uint8)
    keep = np.ones(n, dtype=np.uint8)
    gridpt = np.zeros(n, dtype=np.uint8)
    loc[[2, 3, 7]] = 1
    loc[n - 1] = 1                      # a flag on the last cell is never an insertion (there is no cell after it)
    gridpt[7] = 1
    keep[[5, 9]] = 0
    comp = np.array([1, 0, 1, 0], dtype=np.int32)
    out = _report(dict(loc=loc, keep=keep, gridpt=gridpt, comp=comp), ["u", "v", "w"])
    bar = "#" * 78
    assert out == (f"{bar}\nRefining grid...\nNew points inserted after grid points \n2 3 7     to resolve u w point 7 \n"
                   f"Deleted points at \n5 9 \n{bar}\n")


def test_report_when_nothing_changes():
    n = 5
    flags = dict(loc=np.zeros(n, dtype=np.uint8), keep=np.ones(n, dtype=np.uint8), gridpt=np.zeros(n, dtype=np.uint8),
                 comp=np.zeros(4, dtype=np.int32))
    bar = "#" * 78
    assert _report(flags, ["u"]) == f"{bar}\nNo new points needed\nNo new points deleted\n{bar}\n"


def test_criteria_validation_messages():
    import splitfxm_b200 as sf
    r = sf.Refiner()
    for args in ((1.5, 0.5, 0.5, 0.0), (3.0, 1.5, 0.5, 0.0), (3.0, 0.5, -0.1, 0.0), (3.0, 0.5, 0.5, 0.6)):
        with pytest.raises(sf.SFXM):
            r.set_criteria(*args)
    r.set_criteria(3.0, 0.5, 0.4, 0.1)
    assert (r._ratio, r._slope, r._curve, r._prune) == (3.0, 0.5, 0.4, 0.1)
