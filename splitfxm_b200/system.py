"""System: model + scheme (+ options) holder (reference: splitfxm/system.py:6-70).  ``residuals(d)`` of the
reference is the per-cell Python loop that csrc/sfxm_kernels.cu: residual_kernel replaces; the Simulation
front end calls the kernel directly."""


class System:
    def __init__(self, model, scheme, scheme_opts):
        self._model = model
        self._scheme = scheme
        self._scheme_opts = scheme_opts
