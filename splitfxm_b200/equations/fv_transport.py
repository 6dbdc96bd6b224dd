"""Finite-volume transport equation holder (reference: splitfxm/equations/fv_transport.py:4-59).

Keeps the reference's constructor (F, D, S, dFdU callables) so user models construct unchanged; the
arithmetic of ``residuals`` lives in csrc/sfxm_device.cuh: cell_residual."""


class FVTransportEquation:
    method = "FVM"

    def __init__(self, F, D, S, dFdU):
        self.F = F
        self.D = D
        self.S = S
        self.dFdU = dFdU
