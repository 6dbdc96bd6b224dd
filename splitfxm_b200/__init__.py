"""splitfxm_b200 - B200-native (sm_100a) implementation of SplitFXM's steady-state / time-stepping hot path.

Same public names as the reference package ``splitfxm``: ``Simulation``, ``Domain``, ``Cell``, ``Boundary``,
``Refiner``, ``schemes``, ``bc``, ``initialize``, ``models.advection_diffusion.AdvectionDiffusion`` ...
Importing the package does not need a GPU; any call into the hot path does (no CPU fallback).
"""
from .error import SFXM  # noqa: F401
from .constants import btype, btype_map  # noqa: F401
from .cell import Cell  # noqa: F401
from .boundary import Boundary  # noqa: F401
from .domain import Domain  # noqa: F401
from .schemes import FDSchemes, FVSchemes, default_scheme, stencil_sizes  # noqa: F401
from .model import Model, DeviceModel  # noqa: F401
from .refine import Refiner  # noqa: F401
from .simulation import Simulation  # noqa: F401

__version__ = "0.1.0"
