"""draw(d, label, interior): plot positions against values (reference: splitfxm/visualize.py:5-17).  matplotlib is
imported when the function is called - it is not needed (and not installed in the build image) for anything else."""
from .domain import Domain


def draw(d: Domain, l: str, interior=False):
    from matplotlib.pyplot import plot
    plot(d.positions(interior), d.values(interior), "-o", label=l)
