"""Access to the UNMODIFIED reference (gpavanb1/SplitFXM 0.5.2) installed under baseline/_ref.

Install (build container; the directory is git-ignored but travels to the GPU box with the snapshot):

    cp -r /root/reference /tmp/sfxm_ref_src
    python -m pip install --no-index --no-build-isolation --no-deps --find-links /opt/wheelhouse \
        --target baseline/_ref /tmp/sfxm_ref_src

`--no-deps`: `splitnewton` / `numdifftools` / `matplotlib` are not in the wheelhouse; baseline/stubs holds
stand-ins (numdifftools: forward differences; matplotlib: no-ops; splitnewton: oracle/newton.py, parity
unpinned).  Used only by `bench.py --impl reference` / `cpu_baseline` and the golden-fixture scripts - never by
the product path.
"""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = os.path.join(HERE, "_ref")
STUBS = os.path.join(HERE, "stubs")


def available():
    import glob
    return bool(glob.glob(os.path.join(REF, "splitfxm", "flux*.so"))) and \
        os.path.exists(os.path.join(REF, "splitfxm", "simulation.py"))


def install():
    """Build + install the reference into baseline/_ref from a scratch copy (needs /root/reference)."""
    import shutil
    import subprocess
    import tempfile
    src = "/root/reference"
    if available() or not os.path.isdir(src):
        return available()
    work = os.path.join(tempfile.gettempdir(), "sfxm_ref_src")
    shutil.rmtree(work, ignore_errors=True)
    shutil.copytree(src, work)
    subprocess.check_call([sys.executable, "-m", "pip", "install", "--no-index", "--no-build-isolation", "--no-deps",
                           "--find-links", "/opt/wheelhouse", "--target", REF, work],
                          stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    return available()


def activate():
    """Put the reference (and the stand-ins it needs) first on sys.path; returns the imported `splitfxm` package.
    Call in a process that does not import the product's `splitfxm` alias."""
    if not available():
        raise RuntimeError("baseline/_ref is not installed (see baseline/reference_env.py)")
    for p in (ROOT, STUBS, REF):
        if p in sys.path:
            sys.path.remove(p)
        sys.path.insert(0, p)
    for name in [m for m in sys.modules if m == "splitfxm" or m.startswith("splitfxm.")]:
        del sys.modules[name]
    import splitfxm
    assert os.path.dirname(os.path.abspath(splitfxm.__file__)) == os.path.join(REF, "splitfxm")
    return splitfxm
