"""Model base class (reference: splitfxm/model.py:1-17) plus the device-functor hook.

A model that can run on the GPU carries ``_device = DeviceModel(model_id, params)``: the id selects one of the
compiled functors in csrc/sfxm_device.cuh, ``params`` are its per-member parameters.  A model whose
``_equation`` holds only opaque Python callables cannot be evaluated on the device, and there is no CPU
fallback: Simulation raises SFXM for it."""


class DeviceModel:
    def __init__(self, model_id, params, method):
        self.model_id = int(model_id)
        self.params = [float(p) for p in params]
        self.method = method


class Model:
    def __init__(self):
        self._equation = None
        self._device = None

    def equation(self):
        return self._equation
