"""Advection-diffusion model (reference: splitfxm/models/advection_diffusion.py:9-37)."""
import numpy as np

from .. import _lib
from ..equations.fd_transport import FDTransportEquation
from ..equations.fv_transport import FVTransportEquation
from ..error import SFXM
from ..model import DeviceModel, Model


class AdvectionDiffusion(Model):
    def __init__(self, c, nu, method='FDM'):
        self.c = c
        self.nu = nu
        def F(u): return np.array([self.c * x for x in u])
        def D(u): return np.array([self.nu * x for x in u])
        def S(u): return np.array([0.0])
        def dFdU(u): return np.diag([self.c] * len(u))
        if method == 'FDM':
            self._equation = FDTransportEquation(F, D, S)
        elif method == 'FVM':
            self._equation = FVTransportEquation(F, D, S, dFdU)
        else:
            raise SFXM("Invalid numerical method specified")
        self._device = DeviceModel(_lib.MODEL_AD, [c, nu], method)
