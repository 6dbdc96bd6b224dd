"""Linear advection-diffusion of every variable, u_t + c u_x = nu u_xx
(reference: splitfxm/models/advection_diffusion.py:9-37).

On the device this is ``ModelAD<NV>`` in csrc/sfxm_device.cuh with per-member parameters (c, nu); the host-side
equation object keeps the reference's F / D / S / dFdU attributes for code that inspects a model.
"""
import numpy as np

from .. import _lib
from ..equations.fd_transport import FDTransportEquation
from ..equations.fv_transport import FVTransportEquation
from ..error import SFXM
from ..model import DeviceModel, Model


class AdvectionDiffusion(Model):
    def __init__(self, c, nu, method='FDM'):
        if method not in ("FDM", "FVM"):
            raise SFXM("Invalid numerical method specified")
        self.c, self.nu = c, nu
        self._device = DeviceModel(_lib.MODEL_AD, lambda: [self.c, self.nu], method)
        if method == "FVM":
            self._equation = FVTransportEquation(self.flux, self.diffused, self.source, self.flux_jacobian)
        else:
            self._equation = FDTransportEquation(self.flux, self.diffused, self.source)

    # host twins of the device functor (the kernels never call them)
    def flux(self, u):
        return self.c * np.asarray(u, dtype=np.float64)

    def diffused(self, u):
        return self.nu * np.asarray(u, dtype=np.float64)

    def source(self, u):
        return np.zeros(1)

    def flux_jacobian(self, u):
        return self.c * np.eye(len(u))
