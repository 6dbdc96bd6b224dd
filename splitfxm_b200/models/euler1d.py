"""1-D Euler models of the reference's tutorials, as device models.

Euler1D (primitive rho, u, p):            tutorials/sod_shock/euler1d.py:8-57
Euler1DConservative (rho, rho u, E):      tutorials/conservative_sod_shock/euler1d.py:8-90
"""
import numpy as np

from .. import _lib
from ..equations.fd_transport import FDTransportEquation
from ..equations.fv_transport import FVTransportEquation
from ..error import SFXM
from ..model import DeviceModel, Model


def _equation(method, F, D, S, dFdU):
    if method == 'FDM':
        return FDTransportEquation(F, D, S)
    if method == 'FVM':
        return FVTransportEquation(F, D, S, dFdU)
    raise SFXM("Invalid numerical method specified")


class Euler1D(Model):
    def __init__(self, gamma=1.4, method='FDM'):
        self.gamma = gamma

        def F(U):
            rho, u, p = U
            E = p / (self.gamma - 1) + 0.5 * rho * u**2
            return np.array([rho * u, rho * u**2 + p, u * (E + p)])

        def D(U): return np.zeros_like(U)
        def S(U): return np.zeros_like(U)

        def dFdU(U):
            rho, u, p = U
            return np.array([[u, rho, 0], [u**2, 2 * rho * u, 1],
                             [0.5 * (u**3.), 1.5 * rho * (u**2.) + gamma * p / (gamma - 1.), gamma * u / (gamma - 1.)]])
        self._equation = _equation(method, F, D, S, dFdU)
        self._device = DeviceModel(_lib.MODEL_EULER_PRIM, [gamma], method)


class Euler1DConservative(Model):
    def __init__(self, gamma=1.4, method='FDM', eps=1e-10):
        self.gamma = gamma

        def F(U):
            U1, U2, U3 = U
            return np.array([U2, 0.5 * (3 - gamma) * U2**2 / U1 + (gamma - 1) * U3,
                             gamma * U2 * U3 / U1 - 0.5 * (gamma - 1) * (U2**3 / U1**2)])

        def D(U): return np.zeros_like(U)
        def S(U): return np.zeros_like(U)

        def dFdU(U):
            U1, U2, U3 = U
            D4 = -0.5 * (3 - gamma) * (U2**2 / U1**2)
            D5 = (3 - gamma) * U2 / U1
            D6 = gamma - 1
            D7 = -gamma * U2 * U3 / U1**2 + (gamma - 1) * (U2**3 / U1**3)
            D8 = gamma * U3 / U1 - 1.5 * (gamma - 1) * (U2**2 / U1**2)
            D9 = gamma * U2 / U1
            return np.array([[0, 1, 0], [D4, D5, D6], [D7, D8, D9]])
        self._equation = _equation(method, F, D, S, dFdU)
        self._device = DeviceModel(_lib.MODEL_EULER_CONS, [gamma], method)

    def conservative_to_primitive(self, U):
        U1, U2, U3 = U
        return np.array([U1, U2 / U1, (self.gamma - 1) * (U3 - 0.5 * U2**2 / U1)])
