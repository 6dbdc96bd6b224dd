"""Finite-difference transport equation holder (reference: splitfxm/equations/fd_transport.py:4-51)."""


class FDTransportEquation:
    method = "FDM"

    def __init__(self, F, D, S):
        self.F = F
        self.D = D
        self.S = S
