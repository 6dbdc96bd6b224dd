"""`splitfxm` - drop-in alias of the B200-native package: ``import splitfxm`` and every reference-style submodule import
(``from splitfxm.simulation import Simulation``, ``from splitfxm.models.advection_diffusion import AdvectionDiffusion``,
``from splitfxm.flux import fluxes`` ...) resolve to the modules of ``splitfxm_b200`` - the same module objects, so
classes and the ``SFXM`` exception are shared between the two names."""
import importlib
import pkgutil
import sys

import splitfxm_b200 as _impl
from splitfxm_b200 import *  # noqa: F401,F403
from splitfxm_b200 import __version__  # noqa: F401

for _m in pkgutil.walk_packages(_impl.__path__, _impl.__name__ + "."):
    if ".csrc" in _m.name or _m.name.rsplit(".", 1)[-1].startswith("_"):
        continue
    _mod = importlib.import_module(_m.name)
    sys.modules[__name__ + _m.name[len(_impl.__name__):]] = _mod
    if _m.name.count(".") == 1:
        globals()[_m.name.rsplit(".", 1)[-1]] = _mod
