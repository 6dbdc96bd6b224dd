"""Adaptive mesh refinement front end (reference: splitfxm/refine.py:10-312, a port of Cantera's refine.cpp).

Criteria bookkeeping lives here; the marking / scan / compact passes run on the device (csrc/sfxm_amr.cu).
"""
from .error import SFXM


class Refiner:
    def __init__(self):
        # defaults: refine.py:23-36
        self._ratio = 10.0
        self._slope = 0.8
        self._curve = 0.8
        self._prune = -0.001
        self._npmax = 1000
        self._min_range = 0.01
        self._min_grid = 1e-10

    def set_criteria(self, ratio, slope, curve, prune):
        # refine.py:38-76
        if ratio < 2.0:
            raise SFXM(f"ratio must be greater than 2.0 ({ratio} was specified).")
        elif slope < 0.0 or slope > 1.0:
            raise SFXM(f"slope must be between 0.0 and 1.0 ({slope} was specified).")
        elif curve < 0.0 or curve > 1.0:
            raise SFXM(f"curve must be between 0.0 and 1.0 ({curve} was specified).")
        elif prune > curve or prune > slope:
            raise SFXM(f"prune must be less than 'curve' and 'slope' ({prune} was specified).")
        self._ratio = ratio
        self._slope = slope
        self._curve = curve
        self._prune = prune

    def set_max_points(self, npmax):
        self._npmax = npmax

    def refine(self, d):
        raise SFXM("Refiner.refine: device AMR kernels are not built yet")
