"""The stdout report of a refinement pass (splitfxm_b200/refine.py show_changes) in the reference's format
(refine.py:284-312: banner, inserted indices, reasons, deleted indices), from decision flags alone - host logic, no GPU."""
import contextlib
import io

import numpy as np
import pytest


def _report(flags, names):
    """flags as the device kernels produce them (here CPU tensors) -> the reference's loc / c / keep dictionaries ->
    show_changes(loc, c, keep), the reference's signature (refine.py:284)."""
    import torch
    from splitfxm_b200.refine import Refiner
    loc, c, keep = Refiner._dicts_from_flags({k: torch.from_numpy(v) for k, v in flags.items()}, names)
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        Refiner().show_changes(loc, c, keep)
    return buf.getvalue()


def test_show_changes_takes_the_reference_dictionaries():
    """refine.py:284-312 verbatim behaviour: dict order is print order, only keep == -1 is a deletion."""
    from splitfxm_b200.refine import Refiner
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        Refiner().show_changes({4: 1, 1: 1}, {"v": 1, "point 4": 1}, {0: 1, 3: -1, 2: 1, 7: -1})
    bar = "#" * 78
    assert buf.getvalue() == (f"{bar}\nRefining grid...\nNew points inserted after grid points \n4 1     to resolve v point 4 \n"
                              f"Deleted points at \n3 7 \n{bar}\n")


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
