"""morpho_b200 — B200-native (sm_100a, FP64) implementation of Morpho's functional hot path.

The product is ``lib/libmorpho_cuda.so`` (C ABI in ``include/morpho_cuda.h``); this package is
the thin host-side mirror of the reference's functional API used by tests and benchmarks.
"""
from ._lib import MorphoError, init, lib, set_option  # noqa: F401
from .geometry import Field, Mesh, Selection  # noqa: F401
from .functionals import (Area, AreaEnclosed, EquiElement, Functional, GradSq, Length, LineCurvatureSq, Nematic,  # noqa: F401
                          NematicElectric, NormSq, PairwisePotential, Volume, VolumeEnclosed, LineIntegral, AreaIntegral, VolumeIntegral, ScalarPotential)
from . import integrand  # noqa: F401
from . import meshgen  # noqa: F401
