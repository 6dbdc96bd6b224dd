"""Model base class (reference: splitfxm/model.py:1-17) plus the device-functor hook.

A model that can run on the GPU carries ``_device = DeviceModel(model_id, params)``: the id selects one of the
compiled functors in csrc/sfxm_device.cuh, ``params`` are its per-member parameters.  A model whose
``_equation`` holds only opaque Python callables cannot be evaluated on the device, and there is no CPU
fallback: Simulation raises SFXM for it."""


class DeviceModel:
    """params: a list of numbers, or a callable returning one - evaluated at every solver call, so that a model whose
    attributes are changed after construction (``model.c = 0.3``) is seen with its current values, as in the reference,
    whose F / D read ``self.c`` at call time (models/advection_diffusion.py:28-31)."""

    def __init__(self, model_id, params, method):
        self.model_id = int(model_id)
        self._params = params
        self.method = method

    @property
    def params(self):
        p = self._params() if callable(self._params) else self._params
        return [float(v) for v in p]


class Model:
    def __init__(self):
        self._equation = None
        self._device = None

    def equation(self):
        return self._equation
